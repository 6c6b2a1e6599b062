"""editor_b200 -- B200-native (sm_100a) implementation of the EDITOR hot path behind the reference's model API."""
