__perf_b200_shim__ = True
