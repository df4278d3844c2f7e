_COMP = str.maketrans("ACGTacgtNn", "TGCAtgcaNn")


class Seq(str):
    def __getitem__(self, k):
        r = str.__getitem__(self, k)
        return Seq(r) if isinstance(k, slice) else r

    def reverse_complement(self):
        return Seq(str(self)[::-1].translate(_COMP))

    def lower(self):
        return Seq(str.lower(self))

    def upper(self):
        return Seq(str.upper(self))

    def __add__(self, o):
        return Seq(str(self) + str(o))

    def translate(self, table=None, stop_symbol="*", **kw):
        fwd = getattr(table, "forward_table", {})
        stops = set(getattr(table, "stop_codons", ()))
        s = str(self).upper()
        out = []
        for p in range(0, len(s) - len(s) % 3, 3):
            c = s[p:p + 3]
            out.append(fwd[c] if c in fwd else (stop_symbol if c in stops else "X"))
        return Seq("".join(out))
