"""PAF texts for the tokeniser tests: the golden dataset's map.paf (regenerated with the committed generator and seed) and
hand-made lines for what it does not contain."""
import os
import subprocess
import tempfile

import oracle_ffi

ODD = (b"7\t9000\t10\t800\t+\t3\t900\t0\t790\t700\t800\t60\ttp:A:P\tcm:i:50\tcg:Z:300M2I10M1D478M\n"      # tags before cg:Z:
       b"\n"                                                                                              # empty line
       b"7\t9000\t900\t1700\t-\t4\t900\t5\t795\t700\t800\t0\n"                                           # exactly 12 columns
       b"8\t100\t0\t50\t+\t5\t60\t0\t50\t50\t50\t255\tcg:Z:10=2X5M3N4S\tNM:i:2\n"                         # other operation letters
       b"8\t100\t0\t50\t*\t5\t60\t0\t50\t50\t50\t60\tcg:Z:\n"                                            # empty payload
       b"9\t4294967295\t0\t50\t+\t5\t60\t0\t50\t50\t50\t300\tcg:Z:12\tzz:Z:cg:Z:5M\n"                      # count without a letter, mapq > 255
       b"10\t100\t0\t50\t-x\t5\t60\t0\t50\t50\t50\t60\tcg:Z:7M\tcg:Z:9I")                                 # two cg tags, no final line feed


def golden_paf():
    """map.paf of the golden dataset (tests/golden/make_golden.py: 200 kb genome, 600 reads, seed 7)."""
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run([oracle_ffi.GEN_BIN, tmp, "200000", "600", "8000", "7"], check=True, stdout=subprocess.DEVNULL)
        with open(os.path.join(tmp, "map.paf"), "rb") as f:
            return f.read()


def same_hits(a, b):
    return all(a[k].tobytes() == b[k].tobytes() for k in oracle_ffi.PAF_COLS + ("is_rev", "mapq", "cg_off")) and \
        a["cg_ops"][: int(a["cg_off"][-1])].tobytes() == b["cg_ops"][: int(b["cg_off"][-1])].tobytes()


def line_of(cols, tags=()):
    return b"\t".join(list(cols) + list(tags))


def fuzzed(seed, n_lines=4000):
    """(text, lines): rows assembled from odd tokens — signs, blanks, overflow, empty columns, carriage returns, tags in any order."""
    import numpy as np
    rng = np.random.default_rng(seed)
    toks = [b"0", b"7", b"60", b"255", b"256", b"4294967295", b"4294967296", b"123456789012", b" 9", b"+9", b"-9", b"9x", b"x", b"", b"\r5", b"5\r", b"007", b"- 1"]
    tags = [b"tp:A:P", b"cg:Z:5M", b"cg:Z:", b"cg:Z:3M2I\r", b"cg:Z:10", b"cg:Z:M", b"NM:i:3", b"cg:z:5M", b"xcg:Z:5M", b"cg:Z:4294967296M1I", b"", b"cg:Z:1=2X3N"]
    lines = []
    for _ in range(n_lines):
        cols = [toks[i] for i in rng.integers(0, len(toks), 12)]
        cols[4] = [b"+", b"-", b"", b"-+", b"*"][int(rng.integers(0, 5))]
        lines.append(line_of(cols, [tags[i] for i in rng.integers(0, len(tags), int(rng.integers(0, 4)))]))
        if rng.random() < 0.05:
            lines.append(b"")
    return b"\n".join(lines), lines
