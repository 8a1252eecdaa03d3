"""Empty stand-in for the reference module model_gru (not on the HierTCN hot path). TEST INFRASTRUCTURE."""
