"""Drop-in check: bin/haslr_assemble (GPU) against the reference binary (oracle/_ref/haslr_assemble_ref, built from the
unmodified reference sources) on the same seeded synthetic dataset — every output file byte for byte."""
import filecmp
import os
import subprocess
import tempfile

import pytest

import oracle_ffi

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "bin", "haslr_assemble")

FILES = ["compact_uniq.txt", "backbone.01.init.gfa", "backbone.01.init.stat", "backbone.02.weakEdge.gfa", "backbone.02.weakEdge.stat",
         "backbone.03.tip.gfa", "backbone.03.tip.stat", "backbone.03.tip.log", "backbone.04.simplebubble.gfa", "backbone.04.simplebubble.stat",
         "backbone.04.simplebubble.log", "backbone.05.superbubble.gfa", "backbone.05.superbubble.stat", "backbone.05.superbubble.log",
         "backbone.06.smallbubble.gfa", "backbone.06.smallbubble.stat", "backbone.06.smallbubble.log", "backbone.branching.log",
         "log_coordinate.txt", "log_consensus.txt", "asm.final.fa", "asm.final.ann", "log_asmfinal.txt"]


def run_both(tmp, genome, n_reads, seed, extra=()):
    subprocess.run([oracle_ffi.GEN_BIN, tmp, str(genome), str(n_reads), "8000", str(seed)], check=True, stdout=subprocess.DEVNULL)
    args = ["-t", "1", "-c", "contigs.fa", "-l", "reads.fa", "-m", "map.paf", "--aln-block", "500", "--aln-sim", "0.85", "--edge-sup", "3"]
    with open(os.path.join(tmp, "ref.err"), "w") as e:
        subprocess.run([oracle_ffi.REF_BIN] + args + ["-d", "ref"], cwd=tmp, check=True, stdout=subprocess.DEVNULL, stderr=e)
    with open(os.path.join(tmp, "new.err"), "w") as e:
        r = subprocess.run([BIN] + args + ["-d", "new"] + list(extra), cwd=tmp, stdout=subprocess.DEVNULL, stderr=e)
    if r.returncode != 0:
        with open(os.path.join(tmp, "new.err")) as f:
            raise AssertionError("haslr_assemble failed:\n" + f.read()[-3000:])


@pytest.mark.parametrize("genome,n_reads,seed", [(300000, 900, 3), (2000000, 6000, 5)])
def test_all_output_files_identical(genome, n_reads, seed):
    for p in (oracle_ffi.GEN_BIN, oracle_ffi.REF_BIN, BIN):
        if not os.path.exists(p):
            pytest.skip(f"{p} not built")
    with tempfile.TemporaryDirectory() as tmp:
        run_both(tmp, genome, n_reads, seed)
        bad = [f for f in FILES if not (os.path.exists(os.path.join(tmp, "ref", f)) and os.path.exists(os.path.join(tmp, "new", f))
                                        and filecmp.cmp(os.path.join(tmp, "ref", f), os.path.join(tmp, "new", f), shallow=False))]
        assert not bad, f"files differ from the reference binary's: {bad}"
        assert os.path.getsize(os.path.join(tmp, "new", "asm.final.fa")) > genome // 2


def test_help_exits_zero_like_haslr_py_probe():
    # bin/haslr.py:264-279 runs `<tool> -h` and needs exit status 0
    assert subprocess.run([BIN, "-h"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL).returncode == 0
    assert subprocess.run([BIN], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL).returncode != 0
