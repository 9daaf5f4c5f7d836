"""B200-native stand-in for the reference's `gcn_lib` package (only `gcn_lib.sparse` is on the hot path)."""
