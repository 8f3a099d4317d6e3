"""GRACE import paths (`grace_dl.dist...`) backed by deepreduce_b200.grace."""
