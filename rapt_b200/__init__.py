"""rapt_b200 -- B200-native engine for RAPT's particle-advance hot path (placeholder, filled below)."""
