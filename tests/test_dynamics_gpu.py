"""Device benchmark on a real GPU: the C++ loop times the FULL forward + backward kernel chain of a
transformer block (GEMMs, tcgen05 attention, LayerNorm, reductions) with CUDA events."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_device_benchmark_proxy_is_the_whole_block_and_scales_with_slowdown():
    from skycomputing_b200.ops import native as nat

    ext = nat.ext()
    kw = dict(tokens=2048, hidden=1024, intermediate=4096, iterations=5, warmup=2)
    t_full, free_mib = ext.device_benchmark(**kw)
    t_gemm, _ = ext.device_benchmark(mode=1, **kw)
    assert t_full > 0 and free_mib > 1000
    # forward + backward of the whole block costs 3-5x its four forward GEMMs
    assert 2.5 * t_gemm < t_full < 8 * t_gemm, (t_full, t_gemm)
    # analytic cross-check: 3 x 52.6 GFLOP (fwd + dgrad + wgrad at 2048 tokens) per block; the
    # chain must run at a plausible tensor-core rate (well above any CUDA-core fallback)
    tflops = 5 * 3 * 52.6e9 / t_full / 1e12
    assert 150 < tflops < 2500, tflops
    t_slow, _ = ext.device_benchmark(slowdown=1.0, **kw)      # device-side throttle: 2x
    assert 1.7 * t_full < t_slow < 2.4 * t_full, (t_slow, t_full)
    torch.cuda.synchronize()


def test_device_benchmarker_class_uses_the_native_proxy():
    import skycomputing_b200 as sky

    wm = sky.WorkerManager(first_rank=0)
    wm.load_worker_pool_from_config([dict(name="gpu-0", server_config={}, device=0,
                                          extra_config=dict(slowdown=0, mem_limit=-1, cuda_device=0))])
    shape = dict(tokens=1024, hidden=1024, intermediate=4096)
    full = sky.DeviceBenchmarker(wm, None, model_config=[], iterations=4, warmup=2,
                                 proxy="bert_block", block_shape=shape).benchmark()
    gemm = sky.DeviceBenchmarker(wm, None, model_config=[], iterations=4, warmup=2,
                                 proxy="bert_block",
                                 block_shape=dict(shape, proxy_kernels="gemm_only")).benchmark()
    (name, res), = full.items()
    assert res["time"] > gemm[name]["time"] > 0 and res["avai_mem"] > 0
