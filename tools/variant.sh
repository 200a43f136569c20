#!/bin/bash
# Build an ablation variant of the library and print what ptxas made of the two shipped render_frame instantiations.
#   tools/variant.sh NAME "EXTRA nvcc flags"     ->  yoxel-voxel_b200/libyv_b200_NAME.so   (YV_B200_LIB selects it)
NAME=$1; EXTRA=$2
cd "$(dirname "$0")/../yoxel-voxel_b200/csrc" || exit 1
rm -rf build_$NAME
make -j8 variant NAME=$NAME EXTRA="$EXTRA" > /tmp/variant_$NAME.log 2>&1 || { grep -E "error" /tmp/variant_$NAME.log | head; exit 1; }
python3 - "$NAME" <<'P'
import re, subprocess, sys
txt = open('/tmp/variant_%s.log' % sys.argv[1]).read()
ents = re.findall(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", txt)
dem = subprocess.run(['c++filt'] + [e[0] for e in ents], capture_output=True, text=True).stdout.split('\n')
for d, e in zip(dem, ents):
    if 'render_frame<false, false, 0, false, false, false, false, false, false>' in d or 'render_frame<true, false, 0, false, false, false, false, false, false>' in d:
        print("%-10s stack %s spill st/ld %s/%s regs %s  %s" % (sys.argv[1], e[1], e[2], e[3], e[4], "SEC" if "<true" in d else "primary"))
P
