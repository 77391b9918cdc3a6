"""Allocator parity + optimality (goldens: reference Allocator driven by fake benchmarkers,
SURVEY Appendix A item 5; exact solver vs brute force; hypothesis invariants)."""
import itertools

import pytest
from hypothesis import given, settings, strategies as st

import skycomputing_b200 as sky
from skycomputing_b200 import _core

BIG = [1e9] * 8


def entries(L):
    return [0.0] + [36.507, 34.360, 34.360] * L + [0.067, 0.0]


def counts(b):
    return [b[i + 1] - b[i] for i in range(len(b) - 1)]


GOLD_DYNAMIC = [
    (24, [1] * 8, [10, 10, 10, 9, 9, 9, 9, 9], 352.2),
    (24, [1, 1, 1, 2, 1, 1, 1, 1], [11, 10, 10, 8, 10, 10, 10, 6], 562.6),
    (24, [2, 1, 1, 1, 1, 1, 1, 1], [10, 11, 11, 11, 11, 11, 9, 1], 631.4),
    (24, [1, 1, 1, 1, 1, 1, 1, 2], [11, 10, 10, 10, 10, 10, 10, 4], 352.2),
    (24, [1.0 + 0.2 * i for i in range(8)], [16, 12, 10, 9, 8, 7, 7, 6], 538.6),
    (160, [1, 1, 1, 2, 1, 1, 1, 1], [68, 67, 66, 42, 67, 66, 66, 41], 2946.4),
    (160, [2, 1, 1, 1, 1, 1, 1, 1], [61, 68, 68, 68, 68, 68, 68, 14], 4209.1),
    (160, [1, 1, 1, 1, 1, 1, 1, 2], [68, 66, 66, 65, 64, 64, 64, 26], 2351.5),
    (160, [1.0 + 0.2 * i for i in range(8)], [98, 79, 67, 59, 53, 47, 43, 37], 3403.8),
]
GOLD_EXACT = [
    (24, [1] * 8, 315.7), (24, [1, 1, 1, 2, 1, 1, 1, 1], 352.2), (24, [2, 1, 1, 1, 1, 1, 1, 1], 352.2),
    (24, [1, 1, 1, 1, 1, 1, 1, 2], 352.2), (24, [1.0 + 0.2 * i for i in range(8)], 506.4),
    (160, [1] * 8, 2104.6), (160, [1, 1, 1, 2, 1, 1, 1, 1], 2246.3), (160, [2, 1, 1, 1, 1, 1, 1, 1], 2246.3),
    (160, [1, 1, 1, 1, 1, 1, 1, 2], 2246.3), (160, [1.0 + 0.2 * i for i in range(8)], 3324.3),
]


def test_even_matches_reference():
    assert counts(_core.even_partition(75, 8)) == [10, 10, 10, 9, 9, 9, 9, 9]
    assert counts(_core.even_partition(483, 8)) == [61, 61, 61, 60, 60, 60, 60, 60]
    assert counts(_core.even_partition(15, 1)) == [15]
    assert counts(_core.even_partition(15, 2)) == [8, 7]


@pytest.mark.parametrize("L,t,expect,bott", GOLD_DYNAMIC)
def test_dynamic_compat_reproduces_reference(L, t, expect, bott):
    lf = entries(L)
    r = _core.dynamic_partition(lf, [1.0] * len(lf), t, BIG, compat=True)
    assert counts(r["boundaries"]) == expect
    assert r["bottleneck"] == pytest.approx(bott, abs=0.06)


@pytest.mark.parametrize("L,t,bott", GOLD_EXACT)
def test_exact_fixed_order_matches_bruteforce_goldens(L, t, bott):
    lf = entries(L)
    r = _core.optimal_partition(lf, [1.0] * len(lf), t, BIG, permute=False)
    assert r["bottleneck"] == pytest.approx(bott, abs=0.06)
    rp = _core.optimal_partition(lf, [1.0] * len(lf), t, BIG, permute=True)
    assert rp["bottleneck"] <= r["bottleneck"] + 1e-9


def test_memory_cap_cases():
    lf = entries(24)
    lm = [130.0] + [150.0, 260.0, 160.0] * 24 + [5.0, 1.0]
    r = _core.dynamic_partition(lf, lm, [1] * 8, [1500.0] + [1e9] * 7, compat=True)
    assert counts(r["boundaries"]) == [8, 12, 10, 9, 9, 9, 9, 9]
    assert r["bottleneck"] == pytest.approx(420.9, abs=0.06)
    with pytest.raises(RuntimeError, match="memory allocation failed"):
        _core.dynamic_partition(lf, lm, [1] * 8, [100.0] * 8, compat=True)
    with pytest.raises(RuntimeError, match="memory allocation failed"):
        _core.optimal_partition(lf, lm, [1] * 8, [100.0] * 8)
    # exact solver respects the cap
    r = _core.optimal_partition(lf, lm, [1] * 8, [1500.0] + [1e9] * 7, permute=False)
    b = r["boundaries"]
    assert sum(lm[b[0]:b[1]]) <= 1500.0


def _brute(lf, lm, dt, dm, permute):
    L, D = len(lf), len(dt)
    best = float("inf")
    orders = itertools.permutations(range(D)) if permute else [tuple(range(D))]
    for order in orders:
        for cuts in itertools.combinations(range(1, L), D - 1):
            b = (0,) + cuts + (L,)
            ok, worst = True, 0.0
            for k, d in enumerate(order):
                if sum(lm[b[k]:b[k + 1]]) > dm[d]:
                    ok = False
                    break
                worst = max(worst, dt[d] * sum(lf[b[k]:b[k + 1]]))
            if ok:
                best = min(best, worst)
    return best


@settings(max_examples=40, deadline=None)
@given(st.integers(2, 4).flatmap(lambda D: st.tuples(
    st.lists(st.floats(0.1, 10), min_size=D + 1, max_size=8),
    st.lists(st.floats(0.5, 4), min_size=D, max_size=D),
    st.booleans())))
def test_exact_solver_equals_bruteforce(args):
    lf, dt, permute = args
    lm = [1.0] * len(lf)
    dm = [1e9] * len(dt)
    r = _core.optimal_partition(lf, lm, dt, dm, permute=permute)
    assert r["bottleneck"] == pytest.approx(_brute(lf, lm, dt, dm, permute), rel=1e-6)
    b = r["boundaries"]
    assert b[0] == 0 and b[-1] == len(lf) and all(b[i] < b[i + 1] for i in range(len(b) - 1))
    assert sorted(r["order"]) == list(range(len(dt)))


@settings(max_examples=40, deadline=None)
@given(st.lists(st.floats(0.1, 10), min_size=8, max_size=40),
       st.lists(st.floats(0.5, 4), min_size=2, max_size=6))
def test_exact_le_heuristic_le_even(lf, dt):
    if len(lf) < len(dt):
        return
    lm, dm = [1.0] * len(lf), [1e9] * len(dt)
    even_b = _core.even_partition(len(lf), len(dt))
    even = _core.partition_bottleneck(lf, lm, dt, dm, list(range(len(dt))), even_b)
    heur = _core.dynamic_partition(lf, lm, dt, dm, compat=False)["bottleneck"]
    exact = _core.optimal_partition(lf, lm, dt, dm, permute=False)["bottleneck"]
    assert exact <= heur * (1 + 1e-9) <= even * (1 + 1e-9)


class _FakeDev:
    def __init__(self, times, mems):
        self.t, self.m = times, mems

    def benchmark(self):
        return {f"worker{i}": dict(time=t, avai_mem=m) for i, (t, m) in enumerate(zip(self.t, self.m))}


class _FakeModel:
    def __init__(self, lf, lm):
        self.lf, self.lm = lf, lm

    def benchmark(self):
        return self.lf, self.lm


def _allocator(L, times, **kw):
    lf = entries(L)
    cfg = ([dict(layer_type="BertEmbeddings")] + [dict(layer_type="BertLayer_Head"),
           dict(layer_type="BertLayer_Body"), dict(layer_type="BertLayer_Tail")] * L
           + [dict(layer_type="BertPooler"), dict(layer_type="BertTailForClassification")])
    wm = sky.WorkerManager(first_rank=0)
    wm.load_worker_pool_from_config([dict(name=f"g{i}", server_config={}, extra_config={})
                                     for i in range(len(times))])
    return sky.Allocator(cfg, wm, _FakeModel(lf, [1.0] * len(lf)), _FakeDev(times, [1e9] * len(times)),
                         **kw), cfg


def test_allocator_front_end_assigns_contiguous_spans():
    alloc, cfg = _allocator(24, [1] * 8)
    wm = alloc.even_allocate()
    assert [len(w.model_config) for w in wm.worker_pool] == [10, 10, 10, 9, 9, 9, 9, 9]
    assert sum((w.model_config for w in wm.worker_pool), []) == cfg
    alloc, cfg = _allocator(24, [2, 1, 1, 1, 1, 1, 1, 1], solver="compat")
    wm = alloc.dynamic_allocate()
    assert [len(w.model_config) for w in wm.worker_pool] == [10, 11, 11, 11, 11, 11, 9, 1]
    alloc, cfg = _allocator(24, [2, 1, 1, 1, 1, 1, 1, 1])
    wm = alloc.optimal_allocate()
    assert alloc.last_result["bottleneck"] == pytest.approx(352.2, abs=0.06)
    assert sum((w.model_config for w in wm.worker_pool), []) == cfg
    assert [w.rank for w in wm.worker_pool] == list(range(8))
    assert sorted(w.device for w in wm.worker_pool) == list(range(8))
    spans = [w.layer_range for w in wm.worker_pool]
    assert spans[0][0] == 0 and spans[-1][1] == len(cfg)
    assert all(spans[i][1] == spans[i + 1][0] for i in range(7))


def test_block_granularity_only_cuts_between_blocks():
    alloc, cfg = _allocator(24, [1, 1, 1, 2, 1, 1, 1, 1], granularity="block")
    for fn in (alloc.even_allocate, alloc.dynamic_allocate, alloc.optimal_allocate):
        wm = fn()
        for w in wm.worker_pool[:-1]:
            end = w.layer_range[1]
            assert cfg[end - 1]["layer_type"] in ("BertLayer_Tail", "BertEmbeddings", "BertPooler")
        assert sum((w.model_config for w in wm.worker_pool), []) == cfg
        # re-create the pool for the next strategy (allocation re-ranks it)
        alloc, cfg = _allocator(24, [1, 1, 1, 2, 1, 1, 1, 1], granularity="block")


def test_exact_solver_is_fast_for_paper_scale():
    import time

    lf = entries(160)
    t0 = time.time()
    r = _core.optimal_partition(lf, [1.0] * len(lf), [1, 1, 1, 2, 1, 1, 1, 1], BIG, permute=True)
    assert time.time() - t0 < 1.0
    assert r["bottleneck"] <= 2246.3 + 0.06
    # 64 devices (the paper's cluster size): fixed-order DP + swap search
    t = [1.0 + (i % 7) * 0.3 for i in range(64)]
    r = _core.optimal_partition(lf, [1.0] * len(lf), t, [1e9] * 64, permute=True)
    even = _core.partition_bottleneck(lf, [1.0] * len(lf), t, [1e9] * 64, list(range(64)),
                                      _core.even_partition(len(lf), 64))
    assert r["bottleneck"] < even


@settings(max_examples=30, deadline=None)
@given(st.integers(2, 4), st.integers(2, 3), st.integers(0, 5),
       st.lists(st.floats(0.5, 4.0), min_size=4, max_size=4))
def test_looped_allocation_invariants(D, v, extra, speeds):
    """v x D virtual stages: spans are contiguous in virtual-stage order, cover every layer exactly
    once, none is empty, virtual stage k sits on worker k % D, and a faster device never gets less
    work than a slower one placed on an identical cost profile."""
    import skycomputing_b200 as sky

    L = v * D + extra
    cfg = [dict(layer_type="Linear", in_features=4, out_features=4) for _ in range(L)]

    class DevB:
        def __init__(self, wm):
            self.wm = wm

        def benchmark(self):
            from skycomputing_b200.utils import generate_worker_name

            return {generate_worker_name(w.rank): dict(time=speeds[w.device], avai_mem=1e9)
                    for w in self.wm.worker_pool}

    class ModB:
        def benchmark(self):
            return [1.0] * L, [1.0] * L

    for alloc in ("even", "optimal"):
        wm = sky.WorkerManager(first_rank=0)
        wm.load_worker_pool_from_config([dict(name=f"w{i}", server_config={}, device=i,
                                              extra_config={}) for i in range(D)])
        a = sky.Allocator(cfg, wm, ModB(), DevB(wm), granularity="layer", solver="exact")
        wm = a.allocate(alloc, virtual_stages=v)
        spans = {}
        for d, w in enumerate(wm.worker_pool):
            assert w.device == d and len(w.chunks) == v
            assert len(w.model_config) == sum(e - b for b, e in w.chunks)
            for c, (b, e) in enumerate(w.chunks):
                assert e > b
                spans[c * D + d] = (b, e)
        flat = [spans[k] for k in range(v * D)]
        assert flat[0][0] == 0 and flat[-1][1] == L
        assert all(flat[k][1] == flat[k + 1][0] for k in range(v * D - 1))
        if alloc == "optimal":
            worst = max(speeds[k % D] * (e - b) for k, (b, e) in enumerate(flat))
            even = sky.Allocator(cfg, _fresh_pool(D), granularity="layer").allocate("even", virtual_stages=v)
            ev = {}
            for d, w in enumerate(even.worker_pool):
                for c, (b, e) in enumerate(w.chunks):
                    ev[c * D + d] = e - b
            assert worst <= max(speeds[k % D] * ev[k] for k in range(v * D)) + 1e-9


def _fresh_pool(D):
    import skycomputing_b200 as sky

    wm = sky.WorkerManager(first_rank=0)
    wm.load_worker_pool_from_config([dict(name=f"w{i}", server_config={}, device=i, extra_config={})
                                     for i in range(D)])
    return wm


def test_schedule_planner_reproduces_measured_steps_and_ranks_plans():
    """Closed-form step model vs the measured 8 / 4 / 2-GPU steps (profiles/bench_history.md), and
    the ordering of plans it implies."""
    import skycomputing_b200 as sky

    # per-micro-batch costs of a 3-block stage at 32 sequences, from the committed 8-GPU timeline
    c8 = sky.StageCosts(forward=0.527e-3, backward=0.977e-3, period=1.516e-3)
    p8 = sky.SchedulePlanner(8, c8, blocks_per_stage=3)
    assert p8.step_time(32, 8) == pytest.approx(22.76e-3, rel=0.02)      # measured 22.76 ms
    assert p8.step_time(16, 16) == pytest.approx(23.30e-3, rel=0.05)     # measured 23.30 ms
    assert p8.step_time(8, 32) > 1.4 * p8.step_time(16, 16)              # measured 43.5 vs 30.9 ms
    best = p8.best(256)
    assert best.schedule == "looped" and best.virtual_stages == 3 and best.micro_batch == 32
    assert best.step_seconds < 0.75 * p8.step_time(32, 8)
    plain = p8.best(256, allow_looped=False)
    assert plain.virtual_stages == 1 and plain.micro_batch in (16, 32)
    # derived from the measured single-GPU step (11.72 ms for 32 sequences): 2 and 4 GPUs
    p2 = sky.SchedulePlanner(2, sky.costs_from_single_gpu_step(11.72e-3, 2), blocks_per_stage=12)
    assert p2.step_time(32, 2) == pytest.approx(17.62e-3, rel=0.05)      # measured 17.62 ms
    assert p2.step_time(16, 4) == pytest.approx(18.27e-3, rel=0.08)      # measured 18.27 ms
    p4 = sky.SchedulePlanner(4, sky.costs_from_single_gpu_step(11.72e-3, 4), blocks_per_stage=6)
    assert p4.step_time(16, 8) == pytest.approx(19.92e-3, rel=0.08)      # measured 19.92 ms
    assert p4.best(128).schedule == "looped"
    # one GPU: no pipeline terms, no looped candidates
    p1 = sky.SchedulePlanner(1, sky.costs_from_single_gpu_step(11.72e-3, 1), blocks_per_stage=24)
    assert p1.best(32).virtual_stages == 1 and p1.step_time(32, 1) == pytest.approx(11.72e-3)


def test_looped_allocation_balances_device_sums_not_only_chunks():
    """v chunks per device: the exact solver minimises the most expensive CHUNK; throughput is set
    by the busiest DEVICE (sum of its chunks).  One 2.2x-slow device out of four, 24 equal blocks,
    v = 2: the chunk-level optimum leaves a fast device with 8 blocks while the slow one has 2;
    the refinement moves a block over (max device load 8.0 -> 7.05)."""
    from skycomputing_b200 import _core
    from skycomputing_b200.dynamics.allocator import Allocator

    D, v = 4, 2
    VP = D * v
    uf = [1.05] + [1.0] * 22 + [1.02]
    um = [1.0] * 24
    dt = [1.0, 2.2, 1.0, 1.0]
    dm = [1e9] * D

    def loads(b):
        return [sum(sum(uf[b[k]:b[k + 1]]) * dt[k % D] for k in range(d, VP, D)) for d in range(D)]

    res = _core.optimal_partition(uf, um, [dt[k % D] for k in range(VP)],
                                  [dm[k % D] / v for k in range(VP)], permute=False, min_layers=1,
                                  cut_penalty=[])
    b0 = list(res["boundaries"])
    b1 = Allocator._refine_device_loads(list(b0), uf, um, dt, dm, D)
    assert max(loads(b1)) < max(loads(b0)) - 0.5
    assert max(loads(b1)) == pytest.approx(7.05, abs=0.02)
    assert b1[0] == 0 and b1[-1] == 24 and all(b1[k + 1] > b1[k] for k in range(VP))
    # homogeneous devices: nothing to refine
    res = _core.optimal_partition(uf, um, [1.0] * VP, [dm[0] / v] * VP, permute=False, min_layers=1,
                                  cut_penalty=[])
    assert Allocator._refine_device_loads(list(res["boundaries"]), uf, um, [1.0] * D, dm, D) == \
        list(res["boundaries"])
    # memory caps are respected: a cap that forbids any growth of device 1 keeps its chunks
    tight = [1e9, 2.0, 1e9, 1e9]
    b2 = Allocator._refine_device_loads(list(b0), uf, um, dt, tight, D)
    assert sum(b2[k + 1] - b2[k] for k in range(1, VP, D)) <= 2


def test_comm_aware_allocation_avoids_the_after_body_cut():
    """at layer granularity a cut after BertLayer_Body ships [M, I] + [M, H]
    (5x the bytes of a cut after Head / Tail).  `comm_aware=True` takes the boundary sizes from the
    model benchmarker and charges bytes / link bandwidth per cut; the plain allocator does not care."""
    import skycomputing_b200 as sky
    from skycomputing_b200.models import BertConfig

    c = BertConfig(1000, hidden_size=1024, num_hidden_layers=2, num_attention_heads=16,
                   intermediate_size=4096, max_position_embeddings=128)
    enc = [dict(layer_type="BertLayer_Head", config=c.__dict__),
           dict(layer_type="BertLayer_Body", config=c.__dict__),
           dict(layer_type="BertLayer_Tail", config=c.__dict__)]
    # Head | Body | Tail with a second, lighter block so that the flop-balanced cut of two equal
    # devices falls right after the first Body
    cfg = [dict(layer_type="BertEmbeddings", config=c.__dict__)] + enc + enc[:1]

    class Dev:
        def __init__(self, wm):
            self.wm = wm

        def benchmark(self):
            from skycomputing_b200.utils import generate_worker_name

            return {generate_worker_name(w.rank): dict(time=1.0, avai_mem=1e12)
                    for w in self.wm.worker_pool}

    def split(comm_aware):
        wm = sky.WorkerManager(first_rank=0)
        wm.load_worker_pool_from_config([dict(name=f"w{i}", server_config={}, extra_config={})
                                         for i in range(2)])
        gen = sky.build_data_generator("DataloaderGenerator", generator_cfg=dict(
            dataset_cfg=dict(type="SynthMNLIDataset", num_samples=32, max_seq_length=128),
            dataloader_cfg=dict(batch_size=32)))
        mb = sky.ModelBenchmarker(cfg, gen, device="cpu", analytic=True)
        alloc = sky.Allocator(cfg, wm, mb, Dev(wm), solver="exact", granularity="layer",
                              comm_aware=comm_aware, link_bytes_per_s=340e9,
                              device_flops_per_s=7e14)
        wm = alloc.optimal_allocate(permute=False)
        return [len(w.model_config) for w in wm.worker_pool], mb.last_boundary_bytes

    plain, sizes = split(False)
    aware, _ = split(True)
    # boundary in front of Tail (= after Body) is 5x the one in front of Body (= after Head)
    assert sizes[3] == pytest.approx(5 * sizes[2], rel=0.01) and sizes[0] == 0.0
    assert plain == [3, 2]                    # emb, Head, Body | Tail, Head: flop-balanced
    assert aware in ([2, 3], [4, 1])          # the 80 MiB cut is avoided


def test_refine_looped_partition_properties_random():
    """C++ refine_looped_partition on random instances: valid boundaries, every chunk >= 1 unit,
    memory caps kept, and the (sorted device loads, worst chunk) score never gets worse."""
    import random

    from skycomputing_b200 import _core

    rnd = random.Random(7)
    for _ in range(200):
        D = rnd.randint(2, 5)
        v = rnd.randint(1, 4)
        VP = D * v
        L = rnd.randint(VP, VP + 30)
        uf = [rnd.uniform(0.1, 3.0) for _ in range(L)]
        um = [rnd.uniform(0.5, 2.0) for _ in range(L)]
        dt = [rnd.choice([1.0, 1.0, 1.5, 2.0, 3.0]) for _ in range(D)]
        cuts = sorted(rnd.sample(range(1, L), VP - 1))
        b0 = [0] + cuts + [L]

        def stats(b):
            loads = [sum(sum(uf[b[k]:b[k + 1]]) * dt[k % D] for k in range(d, VP, D)) for d in range(D)]
            mem = [sum(sum(um[b[k]:b[k + 1]]) for k in range(d, VP, D)) for d in range(D)]
            worst = max(sum(uf[b[k]:b[k + 1]]) * dt[k % D] for k in range(VP))
            return loads, mem, worst

        l0, m0, w0 = stats(b0)
        dm = [x * rnd.choice([1.0, 1.2, 10.0]) for x in m0]       # feasible at the start
        b1 = list(_core.refine_looped_partition(uf, um, dt, dm, b0))
        assert b1[0] == 0 and b1[-1] == L and len(b1) == VP + 1
        assert all(b1[k + 1] - b1[k] >= 1 for k in range(VP))
        l1, m1, w1 = stats(b1)
        assert all(m1[d] <= dm[d] + 1e-9 for d in range(D))
        s0 = (tuple(sorted(l0, reverse=True)), w0)
        s1 = (tuple(sorted(l1, reverse=True)), w1)
        assert s1 <= s0
    with pytest.raises(Exception):
        _core.refine_looped_partition([1.0] * 4, [1.0] * 4, [1.0, 1.0], [9.0, 9.0], [0, 2, 3])
