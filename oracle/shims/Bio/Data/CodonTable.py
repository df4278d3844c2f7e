class CodonTable:
    def __init__(self, nucleotide_alphabet=None, protein_alphabet=None, forward_table=None, back_table=None, start_codons=None,
                 stop_codons=None):
        self.forward_table = forward_table or {}
        self.start_codons = start_codons or []
        self.stop_codons = stop_codons or []
