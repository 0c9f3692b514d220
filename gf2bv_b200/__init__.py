"""gf2bv_b200 -- B200-native GF(2) linear-system solver behind gf2bv's LinearSystem API."""
