def GC(seq):
    s = str(seq)
    return 100.0 * sum(s.count(x) for x in "GCgcSs") / len(s) if s else 0.0
