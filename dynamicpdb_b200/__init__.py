"""dynamicpdb_b200: B200-native DFOLDv2 score-network hot path."""
__version__ = "0.1.0"
