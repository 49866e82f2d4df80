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
