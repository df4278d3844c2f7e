class SeqRecord:
    def __init__(self, seq, id="", description="", name=""):
        self.seq, self.id, self.description, self.name = seq, id, description, name

    def __len__(self):
        return len(self.seq)

    def translate(self, table=None, stop_symbol="*", **kw):
        return SeqRecord(self.seq.translate(table=table, stop_symbol=stop_symbol), id=self.id, description=self.description)

    def lower(self):
        return SeqRecord(self.seq.lower(), id=self.id, description=self.description)
