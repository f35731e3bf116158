#!/bin/bash
# usage (on the GPU box, under gpurun): tools/prof_workload.sh <workload> <instances> <tag>
# One `ncu --set full` capture of the workload's step kernel, summarised where it was taken: the text summary and the
# facts bench.py reads back land in gpurun_out/ (the .ncu-rep itself stays on the box unless KEEP_REP=1).
w=$1; inst=$2; tag=$3
rep=/tmp/prof_$w
python bench.py --workload "$w" --steps 6 --warmup 3 --no-cpu --kernel-only > /dev/null 2>&1 || { echo "bench failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 4 -c 1 -f -o $rep \
    python bench.py --workload "$w" --steps 6 --warmup 3 --no-cpu --kernel-only > gpurun_out/ncu_$w.log 2>&1
out=gpurun_out/r02_ncu_${w}_$tag.txt
python profiles/summarize_ncu.py $rep.ncu-rep > $out
python profiles/phase_breakdown.py $rep.ncu-rep >> $out
python profiles/top_stalls.py $rep.ncu-rep >> $out
python profiles/smem_wavefronts.py $rep.ncu-rep >> $out
python profiles/summarize_ncu.py $rep.ncu-rep --json "$w" "$inst" > gpurun_out/r02_ncu_$w.json
[ -n "$KEEP_REP" ] && cp $rep.ncu-rep gpurun_out/
echo "wrote $out"
