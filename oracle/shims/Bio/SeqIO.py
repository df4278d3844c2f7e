from .Seq import Seq
from .SeqRecord import SeqRecord


def parse(path, fmt):
    rid = desc = None
    chunks = []
    with open(path) as fh:
        for line in fh:
            line = line.rstrip("\n")
            if line.startswith(">"):
                if rid is not None:
                    yield SeqRecord(Seq("".join(chunks)), rid, desc)
                desc = line[1:].rstrip()
                rid = desc.split()[0] if desc.split() else ""
                chunks = []
            elif rid is not None:
                chunks.append(line.strip().replace(" ", "").replace("\r", ""))
    if rid is not None:
        yield SeqRecord(Seq("".join(chunks)), rid, desc)


def write(records, path, fmt):
    n = 0
    with open(path, "w") as fh:
        for r in records:
            head = r.description if r.description.startswith(r.id + " ") or r.description == r.id else f"{r.id} {r.description}"
            fh.write(f">{head}\n")
            s = str(r.seq)
            for i in range(0, len(s), 60):
                fh.write(s[i:i + 60] + "\n")
            n += 1
    return n
