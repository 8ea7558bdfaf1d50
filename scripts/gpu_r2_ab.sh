#!/bin/bash
# Round-2 opener (1 GPU, ~15 min): the opt-in kernels were brought up on the CPU kernel emulator in round 1 and have not
# run on hardware.  (1) their parity tests on the GPU, (2) A/B of every option on the configuration it targets.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/gpu_r2_ab.sh'
# Reads: gpurun_out/r2_ab.txt (one line per run), gpurun_out/r2_ab_*.json (full bench lines).
mkdir -p gpurun_out
echo "== opt-in kernel tests on hardware" | tee gpurun_out/r2_ab.txt
SEPGPU_TEST_UNVERIFIED=1 timeout 600 python -m pytest tests/test_gpu_zzz_options.py tests/test_gpu_zz_next.py -m gpu -q 2>&1 | tail -8 | tee -a gpurun_out/r2_ab.txt

run() {   # label, workload, SEPGPU_OPTS, extra bench args
  local tag=$1 wl=$2 opts=$3; shift 3
  SEPGPU_OPTS="$opts" timeout 400 python bench.py --workload $wl --no-cpu --no-e2e "$@" 2>gpurun_out/r2_ab_$tag.err >gpurun_out/r2_ab_$tag.json
  python - "$tag" "$opts" <<'PY' | tee -a gpurun_out/r2_ab.txt
import json, sys
tag, opts = sys.argv[1], sys.argv[2]
try:
    d = json.load(open(f"gpurun_out/r2_ab_{tag}.json"))
    km = {k: round(v["total_ms"] / max(v["launches"], 1), 4) for k, v in d.get("kernel_ms", {}).items()}
    print(f"{tag:28s} opts=[{opts}] value={d['value']:.4e} ms/step={d['ms_per_step']:.4f} kernels(avg ms)={km} "
          f"epot/N={d['config'].get('epot_per_atom')} rebuilds={d['config'].get('list_rebuilds_in_timed_region')}")
except Exception as e:
    print(f"{tag:28s} opts=[{opts}] FAILED: {e}")
    print(open(f"gpurun_out/r2_ab_{tag}.err").read()[-600:])
PY
}

echo "== C1: 1M-atom Lennard-Jones, NVT" | tee -a gpurun_out/r2_ab.txt
S="--steps 600 --warmup 200"
run lj_default        lj "" $S
run lj_prune          lj "build_prune=1" $S
run lj_cellorder      lj "cell_order=1" $S
run lj_pairtile       lj "pair_tile=1" $S
run lj_pairtile_co    lj "pair_tile=1,cell_order=1" $S
run lj_all            lj "pair_tile=1,cell_order=1,build_prune=1" $S
run lj_all_ctas4      lj "pair_tile=1,cell_order=1,build_prune=1,pt_ctas=4" $S
run lj_all_ctas6      lj "pair_tile=1,cell_order=1,build_prune=1,pt_ctas=6" $S
run lj_all_grid2368   lj "pair_tile=1,cell_order=1,build_prune=1,force_grid=2368" $S
run lj_default_g2368  lj "force_grid=2368" $S
run lj_finmulti       lj "fin_multi=1" $S
run lj_stepfold       lj "step_fold=1" $S
run lj_stepfold_multi lj "step_fold=1,fin_multi=1" $S
run lj_everything     lj "pair_tile=1,cell_order=1,build_prune=1,step_fold=1,fin_multi=1" $S
echo "== C2: butane 864k atoms" | tee -a gpurun_out/r2_ab.txt
S="--steps 300 --warmup 50"
run butane_default    butane "" $S
run butane_all        butane "pair_tile=1,cell_order=1,build_prune=1" $S
run butane_finmulti   butane "fin_multi=1" $S
run butane_stepfold   butane "step_fold=1,fin_multi=1" $S
echo "== C3: water 1.12M atoms" | tee -a gpurun_out/r2_ab.txt
run water_default     water "" $S
run water_coul2       water "coulomb_kernel=2" $S
run water_sublist     water "typed_sublist=1" $S
run water_all         water "coulomb_kernel=2,typed_sublist=1,build_prune=1" $S
run water_all_ctas4   water "coulomb_kernel=2,typed_sublist=1,build_prune=1,coul2_ctas=4" $S
run water_all_ctas6   water "coulomb_kernel=2,typed_sublist=1,build_prune=1,coul2_ctas=6" $S
echo "== whole default suite with every option on (SEPGPU_OPTS reaches every context the tests create)" | tee -a gpurun_out/r2_ab.txt
SEPGPU_OPTS="step_fold=1,fin_multi=1,build_prune=1,cell_order=1,pair_tile=1,coulomb_kernel=2,typed_sublist=1" timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee -a gpurun_out/r2_ab.txt
