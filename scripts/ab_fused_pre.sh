# A/B of the fused per-edge product on sub-wave grids (PSI_NO_FUSED_PRE=1 restores the pre-pass launch)
for sw in on off; do
  if [ $sw = off ]; then export PSI_NO_FUSED_PRE=1; else unset PSI_NO_FUSED_PRE; fi
  for wl in c0 c2; do
    timeout 200 python bench.py --workload $wl --no-cpu-baseline --steps 8 --warmup 3 > gpurun_out/r03_${wl}_$sw.json 2> gpurun_out/r03_${wl}_$sw.err
    python - <<PY
import json
d=json.loads(open("gpurun_out/r03_${wl}_$sw.json").read().strip().splitlines()[-1])
print("$wl $sw", d["value"], d["ms_per_step"], d.get("solver_steps_per_step"), {k:(v["avg_us"]) for k,v in (d.get("kernels") or {}).items()}, d["roofline"].get("avg_launch_us"))
PY
  done
done
