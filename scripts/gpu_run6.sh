mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_parity.py -q --timeout 300 2>&1 | tail -12 > gpurun_out/r6_tests.log
tail -6 gpurun_out/r6_tests.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r6_bench.log 2>&1
python - <<'PY'
import json
l=open('gpurun_out/r6_bench.log').read().strip().splitlines()[-1]
d=json.loads(l)
print(d['ms_per_step'], d['value'], d['e2e'], d['roofline']['achieved'])
print(d['stage_ms_per_step'])
PY
python - <<'PY'
# s4 variant diagnostics: per-tap error in bf16
import dataclasses, torch, sys
sys.path.insert(0, '.')
from oracle import dgsct_oracle as O
from dgsct_b200.adapter import adapter_forward_raw
for variant, kw in [("s4", dict(use_bn=False, use_gate=False)), ("ave", {})]:
    cfg = O.make_config(variant, 768, 144, 384, 256, **kw)
    params = O.random_params(cfg, seed=7); x, other = O.random_inputs(cfg, 10, seed=11)
    xq, oq = x.bfloat16(), other.bfloat16()
    p = {k: v.cuda() for k, v in params.items()}
    out, smap, tmap, taps = adapter_forward_raw(dataclasses.asdict(cfg), p, xq.cuda(), oq.cuda(), want_taps=True)
    ro, rs, _, _, inter = O.adapter_forward(cfg, params, xq.float(), oq.float(), want=True)
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
    print(variant, {k: f"{rel(taps[k].cpu(), inter[k]):.2e}" for k in ("V", "Tp", "X1", "m", "s", "Y", "Z")}, "out", f"{rel(out.float().cpu(), ro):.2e}")
PY
