"""B200-native multi-view 3D reconstruction hot path."""
