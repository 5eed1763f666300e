#!/bin/bash
# BASELINE config 2 (synthetic 10 Mb, 50k reads x 8 kb) end to end: reference binary (CPU, -t all cores) vs bin/haslr_assemble (GPU).
set -e
T=${1:-/tmp/cfg2}; mkdir -p $T; cd $T
[ -f map.paf ] || /root/repo/oracle/_ref/gen_synth . 10000000 50000 8000 1 > gen.log
NT=$(nproc)
ARGS="-c contigs.fa -l reads.fa -m map.paf --aln-block 500 --aln-sim 0.85 --edge-sup 3"
rm -rf new; [ -n "$SKIP_REF" ] || rm -rf ref     # SKIP_REF keeps an earlier reference run to compare against
[ -n "$SKIP_REF" ] || ( time /root/repo/oracle/_ref/haslr_assemble_ref -t $NT $ARGS -d ref > ref.out 2> ref.err ) 2> ref.time
( time /root/repo/bin/haslr_assemble -t $NT $ARGS -d new > new.out 2> new.err ) 2> new.time
[ -n "$SKIP_REF" ] || echo "== reference ($NT threads)"; [ -n "$SKIP_REF" ] || grep -E "^\[NOTE\]|elapsed" ref.err | paste - - | sed 's/\[NOTE\] //' | cut -c1-150; [ -n "$SKIP_REF" ] || cat ref.time
echo "== haslr_b200"; grep -E "^\[NOTE\]|elapsed" new.err | paste - - | sed 's/\[NOTE\] //' | cut -c1-150; cat new.time
if [ -d ref ]; then for f in compact_uniq.txt backbone.01.init.gfa backbone.02.weakEdge.gfa backbone.06.smallbubble.gfa log_coordinate.txt log_consensus.txt asm.final.fa asm.final.ann; do cmp -s ref/$f new/$f && echo "same $f" || echo "DIFF $f"; done; else echo "(no reference run in $T/ref: nothing compared)"; fi
grep -c ">" new/asm.final.fa; grep -h "segments\|Mbases\|^\[poa\]" new.err | head -60
