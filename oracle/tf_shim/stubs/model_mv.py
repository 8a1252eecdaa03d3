"""Empty stand-in for the reference module model_mv (not on the HierTCN hot path). TEST INFRASTRUCTURE."""
