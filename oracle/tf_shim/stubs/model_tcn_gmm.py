"""Empty stand-in for the reference module model_tcn_gmm (not on the HierTCN hot path). TEST INFRASTRUCTURE."""
