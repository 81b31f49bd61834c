#!/bin/bash
# Round-2 profile set (run under gpurun): launch list of the C3 bench, ncu --set full of the solver, prep and WFA kernels.
tag=${1:-r2f}
profiles/launch_list.sh $tag > /dev/null
ncu --set full --clock-control none --import-source on -k regex:astar_solve_kernel -s 4 -c 1 -o gpurun_out/${tag}_astar_c3 python profiles/ncu_astar_c3.py 10000 1 > gpurun_out/${tag}_ncu_astar.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:astar_prep_kernel -s 1 -c 1 -o gpurun_out/${tag}_prep_c3 python profiles/ncu_astar_c3.py 10000 1 > gpurun_out/${tag}_ncu_prep.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wfa_align_kernel -s 1 -c 1 -o gpurun_out/${tag}_wfa_c4 python profiles/ncu_wfa_c4.py 100 > gpurun_out/${tag}_ncu_wfa.log 2>&1
tail -2 gpurun_out/${tag}_ncu_astar.log gpurun_out/${tag}_ncu_prep.log gpurun_out/${tag}_ncu_wfa.log
cat gpurun_out/${tag}_launch_shares.txt | head -5
