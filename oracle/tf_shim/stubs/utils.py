"""Empty stand-in for the reference module utils (not on the HierTCN hot path). TEST INFRASTRUCTURE."""
