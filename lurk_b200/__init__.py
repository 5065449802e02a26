"""lurk_b200 -- B200-native STARK proving hot path for Lurk."""
