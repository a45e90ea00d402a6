"""Stand-in for xformers 0.0.22 (not installed)."""
