"""GPU tier, NOT YET RUN ON HARDWARE: kernels written after round 1's GPU budget was spent.
Their oracles are pinned to the compiled reference on CPU (tests/test_oracle_more.py); these tests compare the CUDA path with
the oracle and the reference-generated fixtures and are skipped unless B200_RUN_UNVERIFIED=1 (tests/conftest.py) — to be run,
fixed if needed and un-gated at the start of the next round."""
import os

import numpy as np
import pytest

import cpulibs as cl

pytestmark = [pytest.mark.gpu, pytest.mark.hw_unverified]
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def on_stream(device):
    import torch
    return torch.cuda.stream(torch.cuda.ExternalStream(device.stream))


# ---------------------------------------------------------------------------------------------- mpegvideo inverse quantisers
def gpu_unquant(device, variant, cfg, blocks, blk_n, q, last):
    import torch
    from ffmpeg_b200 import mpegvideo as mv
    from ffmpeg_b200._lib import MpvUnquant
    p = cl.unquant_params(struct=MpvUnquant, **cfg)
    with on_stream(device):
        db = torch.from_numpy(np.ascontiguousarray(blocks)).cuda()
        dn = torch.from_numpy(blk_n).cuda() if blk_n is not None else None
        dq, dl = torch.from_numpy(q).cuda(), torch.from_numpy(last).cuda()
        mv.unquantize_batch_device(device, variant, p, db, blocks.shape[0], dn, dq, dl)
        device.sync()
        return db.cpu().numpy()


def test_unquant_golden_and_oracle(device):
    g = np.load(os.path.join(G, "unquant.npz"))
    for variant in range(7):
        for seed in (11, 12):
            cfg, blocks, blk_n, q, last = cl.unquant_case(seed * 7 + variant, variant, nblocks=48)
            out = gpu_unquant(device, variant, cfg, blocks, blk_n, q, last)
            assert np.array_equal(out, g[f"v{variant}_s{seed}"]), (cl.UNQUANT_VARIANTS[variant], seed)
        for seed in range(6):
            cfg, blocks, blk_n, q, last = cl.unquant_case(300 + seed * 7 + variant, variant, nblocks=5000 + seed)
            use_n = blk_n if seed & 1 else None
            out = gpu_unquant(device, variant, cfg, blocks, use_n, q, last)
            assert np.array_equal(out, cl.orc_unquant(variant, cfg, blocks, use_n, q, last)), (cl.UNQUANT_VARIANTS[variant], seed)


def test_unquant_errors(device):
    import torch
    import ffmpeg_b200 as fb
    from ffmpeg_b200 import mpegvideo as mv
    from ffmpeg_b200._lib import MpvUnquant
    cfg, blocks, blk_n, q, last = cl.unquant_case(1, 0, nblocks=8)
    p = cl.unquant_params(struct=MpvUnquant, **cfg)
    p.permutated[5] = p.permutated[6]                                  # not a permutation
    with on_stream(device):
        db, dq, dl = torch.from_numpy(blocks).cuda(), torch.from_numpy(q).cuda(), torch.from_numpy(last).cuda()
        with pytest.raises(fb.B200Error):
            mv.unquantize_batch_device(device, 0, p, db, 8, None, dq, dl)
        p = cl.unquant_params(struct=MpvUnquant, **cfg)
        with pytest.raises(fb.B200Error):
            mv.unquantize_batch_device(device, 7, p, db, 8, None, dq, dl)
        assert mv.unquantize_batch_device(device, 0, p, db, 0, None, dq, dl) == 0
