"""flamingpy_b200: B200-native backend for FlamingPy's Monte-Carlo EC trial path."""

__version__ = "0.1.0"
