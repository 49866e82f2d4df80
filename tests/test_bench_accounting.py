"""The roofline denominators bench.py reports must be SURVEY.md section 8(d)'s algorithmic figures."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_layer_list_is_the_48_calls_of_one_ave_forward():
    b = _bench()
    calls = b.layer_list()
    assert len(calls) == 48
    assert [c[0] for c in calls[:4]] == ["aud", "vis", "aud", "vis"]            # p1 then p2 of the first paired layer
    assert sorted(set(c[3] for c in calls)) == [0, 1, 2, 3]
    assert sum(1 for c in calls if c[3] == 2) == 24                             # six stage-2 layers x 2 directions x 2 positions


def test_algorithmic_bytes_and_flops_match_the_survey():
    b = _bench()
    for BT, gb, tf in ((10, 0.854, 0.429), (80, 4.193, 3.432)):                 # bf16, L/192 shapes
        tot_b = sum(b.alg_bytes(*b.dims(k, s), BT) for k, _, _, s in b.layer_list())
        tot_f = sum(b.alg_flops(*b.dims(k, s), BT) for k, _, _, s in b.layer_list())
        assert abs(tot_b / 1e9 - gb) < 0.01 * gb, tot_b
        assert abs(tot_f / 1e12 - tf) < 0.01 * tf, tot_f


def test_adapter_parameter_count():
    b = _bench()
    import torch
    gen = torch.Generator().manual_seed(0)
    total = 0
    for kind, _, _, s in b.layer_list():
        C, N, Co, No = b.dims(kind, s)
        p = b.make_params(C, N, Co, No, gen)
        total += sum(v.numel() for k, v in p.items() if "running" not in k)
    # 188.5 M over the 48 adapters (gate_tk, one scalar per adapter, is not part of the engine's table)
    assert abs(total - 188.5e6) < 0.2e6, total


def test_workloads_cover_the_baseline_configs():
    """BASELINE.json configs: ave (1-3), s4 (4: B=4 clips x 5 frames, bicubic resampler), avqa_t60 (5: T=60, K=2, g=4), b224."""
    b = _bench()
    assert set(b.WORKLOADS) == {"ave", "b224", "s4", "avqa_t60"}
    assert (b.WORKLOADS["s4"]["B"], b.WORKLOADS["s4"]["T"], b.WORKLOADS["s4"]["cfg"]["resampler"]) == (4, 5, "bicubic")
    assert (b.WORKLOADS["avqa_t60"]["T"], b.WORKLOADS["avqa_t60"]["cfg"]["K"], b.WORKLOADS["avqa_t60"]["cfg"]["g"]) == (60, 2, 4)
    assert b.WORKLOADS["b224"]["vis"] == [(128, 3136), (256, 784), (512, 196), (1024, 49)]
    # the S4 stencil needs square token grids on both sides
    import math
    for (C, N) in b.WORKLOADS["s4"]["vis"] + b.AUD:
        assert math.isqrt(N) ** 2 == N


def test_stage_work_models_match_the_survey_totals():
    b = _bench()
    b._WL["name"] = "ave"
    w = b.stage_work(80, b.ADAPTER_CFG)
    # reference-order conv flops: 2.04 TFLOP of the resampler's 2.495 (the fc product is the rest); the re-associated
    # product contracts over min(C, C') columns, so the executed conv flops are lower
    ref_order = sum(2 * 80 * b.dims(k, s)[1] * b.dims(k, s)[3] * b.dims(k, s)[2] for k, _, _, s in b.layer_list())
    assert abs(ref_order / 1e12 - 2.04) < 0.02
    assert 0.55 * ref_order < w["resample_conv"][0] < ref_order and w["resample_conv"][1] == "tensor"
    assert w["bwd_dwres"][0] == w["resample_conv"][0] == w["bwd_dother"][0]
    # row kernels: bytes = rows x C x bytes per element moved
    rows_c = sum(80 * b.dims(k, s)[1] * b.dims(k, s)[0] for k, _, _, s in b.layer_list())
    assert w["gate_ln"] == (rows_c * 4, "hbm", w["gate_ln"][2]) and w["bwd_out"][0] == rows_c * 8
    # fused latent-token attention: the token side reads V once (+ the 32 x C latent tokens in / out per frame), the stream side
    # reads X and writes X' and Q, the backward reads dX', X, V and writes dX, dV
    kc = sum(80 * 32 * b.dims(k, s)[0] for k, _, _, s in b.layer_list())
    rows = sum(80 * b.dims(k, s)[1] for k, _, _, s in b.layer_list())
    assert w["tok_update"][0] == rows_c * 2 + kc * 6 and w["tok_update"][1] == "hbm"
    assert w["x_update"][0] == rows_c * 4 + rows * 2 * 32 + kc * 2
    assert w["bwd_token"][0] == rows_c * 10 + rows * 2 * 32 + kc * 16
    assert w["tok_update"][0] + w["x_update"][0] < 0.3 * 4.193e9 * 4          # a fraction of the forward's 4.19 GB, by construction
    b._WL["name"] = "s4"
    ws = b.stage_work(20, b.WORKLOADS["s4"]["cfg"])
    assert "resample_conv" not in ws and "bwd_dw" in ws          # the stencil is a gather, not a GEMM
    b._WL["name"] = "ave"
