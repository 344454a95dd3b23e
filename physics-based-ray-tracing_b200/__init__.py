"""B200-native ray-traced acquisition / path-tracing core (imported as ``prt_b200``)."""
__version__ = "0.1.0"
