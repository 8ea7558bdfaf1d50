#!/bin/bash
# Profile set for the opt-in kernels (1 GPU): launch list + full captures with every Lennard-Jones option on, and the
# two water kernels.  Run AFTER scripts/gpu_r2_ab.sh has shown them green on hardware.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/gpu_profile_optin.sh'
mkdir -p gpurun_out
export SEPGPU_OPTS="pair_tile=1,cell_order=1,build_prune=1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_optin.csv python bench.py --steps 20 --warmup 10 --no-cpu --no-e2e > gpurun_out/launches_optin.log 2>&1
B="python bench.py --steps 10 --warmup 100 --no-cpu --no-e2e"
ncu --set full --clock-control none --import-source on -k regex:k_lj_pairtile -s 100 -c 1 -f -o gpurun_out/prof_k_lj_pairtile $B > gpurun_out/prof_k_lj_pairtile.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_build_tile2 -s 15 -c 1 -f -o gpurun_out/prof_k_build_tile2_pair $B > gpurun_out/prof_k_build_tile2_pair.log 2>&1
export SEPGPU_OPTS="coulomb_kernel=2,typed_sublist=1,build_prune=1"
W="python bench.py --workload water --steps 10 --warmup 30 --no-cpu --no-e2e"
ncu --set full --clock-control none --import-source on -k regex:k_coulomb_list2 -s 30 -c 1 -f -o gpurun_out/prof_k_coulomb_list2 $W > gpurun_out/prof_k_coulomb_list2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_water_optin.csv $W > gpurun_out/launches_water_optin.log 2>&1
ls -la gpurun_out/ | tail -12
