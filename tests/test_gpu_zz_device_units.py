"""Kernel-level device tests: the standalone programs in tests/cuda/ (built by scripts/build_cuda_tests.sh, which
__graft_entry__.build() runs) check conv_tc_kernel / pair_tc_kernel directly against a naive CUDA evaluation of the same
convolution, outside the layer table and the C-ABI.  The watchdog flavour is used: every mbarrier wait is clock-bounded and
traps with a call-site code instead of hanging.  Each configuration runs in its own process (a trap poisons the context)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "build")

CONV_DIRECT = [0, 1, 2, 3, 4, 5, 6, 10, 11, 12, 13, 14]          # test_conv_tc.cu: small shapes, per-thread stores
CONV_STAGED = [0, 1, 2, 3, 4, 5, 6, 7, 10, 11, 12, 13, 14, 25]   # ... and the staged TMA-store epilogue
PAIR_SMALL = list(range(9))                                       # test_pair_tc.cu: every non-perf configuration
RB_SMALL = list(range(9))                                         # test_rb_tc.cu: every non-perf configuration
TZ_SMALL = list(range(12))                                        # test_pair_tz.cu: every non-perf configuration


def _run(binary, cfg, env_extra):
    path = os.path.join(BUILD, binary)
    if not os.path.exists(path):
        pytest.fail("%s not built: __graft_entry__.build() / scripts/build_cuda_tests.sh must run before the GPU suite "
                    "(a missing binary must not make these tests vanish)" % binary)
    env = dict(os.environ)
    env.update(env_extra)
    r = subprocess.run([path, str(cfg), "1"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "PASS" in r.stdout, (r.stdout[-1500:], r.stderr[-500:])


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,staged", [(c, "0") for c in CONV_DIRECT] + [(c, "1") for c in CONV_STAGED])
def test_conv_tc_kernel_against_naive_cuda(cfg, staged):
    """Dilated Conv1d / polyphase ConvTranspose1d shapes of the generator (direct and staged TMA-store epilogue)."""
    _run("test_conv_tc_wd", cfg, {"E2E_CONV_STAGED": staged})


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", PAIR_SMALL)
def test_pair_tc_kernel_against_naive_cuda(cfg):
    """Fused x + c2(lrelu(c1(lrelu(x)))) for C = 32 / 64 / 128, one-CTA and CTA-pair forms, ragged tails."""
    _run("test_pair_tc_wd", cfg, {})


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", RB_SMALL)
def test_rb_tc_kernel_against_naive_cuda(cfg):
    """Fused whole ResBlock1 (three pairs chained, residual stream in TMEM) for C = 32 / 64 / 128, k = 3 / 5 / 7,
    one to three pairs, ragged tails, running-sum / divide epilogues."""
    _run("test_rb_tc_wd", cfg, {})


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", TZ_SMALL)
def test_pair_tz_kernel_against_naive_cuda(cfg):
    """Fused pair of the C = 32 stage with four time steps per GEMM row (sliding-window weights, N = 128 MMAs; dilated c1
    as N = 32 MMAs): k = 3 / 5 / 7 / 11, d = 1 / 2 / 3 / 5, ragged tails, running sums in the natural and the tiled8
    layout, odd unit counts per CTA."""
    _run("test_pair_tz_wd", cfg, {})
