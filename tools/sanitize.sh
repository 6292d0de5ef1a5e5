#!/bin/bash
# compute-sanitizer pass over one small parity case per kernel family (run on the GPU box):
#   tools/sanitize.sh [racecheck|memcheck|synccheck|initcheck ...]
# Logs go to gpurun_out/sanitizer_<tool>.log; the summaries are copied to profiles/ by hand.
# Kernel families covered by the selected tests:
#   dg_tendency_kernel Euler / AUX / VISC / SRCX, dg_gradient_kernel (+HYPER), hyper_divergence_kernel,
#   hyper_flux_kernel, hb_filter / hb_gradient / hb_column / hb_tendency, tracer_gradient / tracer_tendency,
#   filter_kernel, courant_kernel, lsrk_update / scale / pack / unpack kernels.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TOOLS=${@:-racecheck memcheck}
SEL='test_vortex_tendency_and_step and rusanov or test_baroclinic_wave_cubed_sphere and roe or test_viscous_box_second_order_path and turbulence0 or test_held_suarez_forcing_and_sponge and turbulence0 or test_dry_biharmonic_hyperdiffusion and sphere-turbulence1 or test_ocean_hbmodel_tendency_and_steps or test_tracers_constant_viscosity_and_inviscid and constant_kinematic or test_filters_apply and indices-every or test_per_step_filter_in_fused_stepper or test_courant_numbers or test_vortex_float32'
for tool in $TOOLS; do
  log=gpurun_out/sanitizer_${tool}.log
  echo "== compute-sanitizer --tool $tool" > "$log"
  extra=""
  [ "$tool" = racecheck ] && extra="--racecheck-report all"
  [ "$tool" = memcheck ] && extra="--leak-check no"
  timeout 1500 compute-sanitizer --tool "$tool" $extra --print-limit 50 --launch-timeout 0 \
    --target-processes application-only --log-file "${log}.raw" \
    python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$SEL" >> "$log" 2>&1
  echo "exit code: $?" >> "$log"
  echo "== sanitizer summary" >> "$log"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|Error" "${log}.raw" | sort | uniq -c | head -60 >> "$log"
  head -c 200000 "${log}.raw" > "${log}.raw.head"; rm -f "${log}.raw"
done
