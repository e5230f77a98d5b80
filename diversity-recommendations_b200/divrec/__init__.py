"""divrec (B200-native hot path) -- same import surface as the reference package for the hot path."""
