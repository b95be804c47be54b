"""B200-native photon-packet propagation engine behind the Hyperion model-file interface."""
__version__ = "0.1.0"
