"""Empty stand-in for the reference module model_gru_gmm (not on the HierTCN hot path). TEST INFRASTRUCTURE."""
