"""Stand-in module; never called by the hot path."""
