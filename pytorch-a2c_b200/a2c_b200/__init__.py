"""a2c_b200 -- B200-native A2C rollout post-processing + update hot path."""
