ncu --set full --import-source on --clock-control none -k regex:anchor_hidden_tc2 -s 6 -c 1 -o gpurun_out/r1_anchor_tc2_v3 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-graph > gpurun_out/ncu_a3.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:pairwise_tc -s 3 -c 1 -o gpurun_out/r1_pairwise_tc_v3 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-graph > gpurun_out/ncu_p3.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:project_tc -s 3 -c 1 -o gpurun_out/r1_project_tc python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-graph > gpurun_out/ncu_j3.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:anchor_hidden_tc2 -s 2 -c 1 -o gpurun_out/r1_shared_conv python tools/bench_shared_conv.py --maps 2 --iters 1 > gpurun_out/ncu_c3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base function -k regex:"gather|anchor|project|pairwise|aff_|softmax" -s 33 -c 44 --csv --log-file gpurun_out/r1_launches_v4.csv python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e --no-graph > gpurun_out/ncu_l4.log 2>&1
ls -la gpurun_out/*.ncu-rep
