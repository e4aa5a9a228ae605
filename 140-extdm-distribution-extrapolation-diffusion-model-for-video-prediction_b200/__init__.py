"""ExtDM sampling hot path, B200-native."""
