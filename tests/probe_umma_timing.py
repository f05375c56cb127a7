"""Run on the GPU box: tensor-pipe time per tcgen05.mma for the operand shapes the fp16-pair engine uses (one CTA, one issuing thread that
advances the descriptors in registers; SM cycles from the first issue to the completion of the commit, and of the issue loop alone;
csrc/umma_probe.cu umma_timing).  Not a pytest file:
    python tests/probe_umma_timing.py  > gpurun_out/umma_timing.txt"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from probe_umma import idesc, lib, sdesc          # noqa: E402

N_MMA = 640          # a multiple of every (period x accumulators) body: 4, 8, 16, 20
A_OFF, B_OFF, SZ = 0, 65536, 131072


def timed(name, ts, id_, a0, a_step, a_per, b0, b_step, b_per, d_stride, n_acc, n=N_MMA, n_warps=1, uniform=0):
    l = lib()
    l.umma_timing.restype = C.c_int
    l.umma_timing.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint64, C.c_uint32, C.c_int, C.c_uint64, C.c_uint32, C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_void_p]
    out = (C.c_longlong * 2)()
    rc = l.umma_timing(SZ, n, ts, id_, a0, a_step, a_per, b0, b_step, b_per, d_stride, n_acc, n_warps, uniform, C.cast(out, C.c_void_p))
    if rc:
        print('%-74s ERROR %s' % (name, l.umma_probe_error().decode()))
        return
    tot = n * n_warps
    print('%-86s %8d cycles = %6.1f per MMA   (issue loop %6.1f per MMA and warp)' % (name, out[0], out[0] / tot, out[1] / n), flush=True)


def kmajor(name, N, n_acc, M=128, ts=False, **kw):
    # K-steps of 16 = two 8-element chunks: A chunk stride 2,048 B (128 rows x 16 B), B chunk stride N x 16 B; four K-steps cycle
    timed(name, 1 if ts else 0, idesc(0, M, N), 8 if ts else sdesc(A_OFF, 2048, 128), 8 if ts else 4096 >> 4, 4,
          sdesc(B_OFF, N * 16, 128), (2 * N * 16) >> 4, 4, max(64, N) if kw.get('n_warps', 1) == 1 else 32, n_acc, **kw)


def mnmajor(name, M, N, n_acc=1):
    # both operands [chunk of 8 units][128 points][8]: K-steps of 16 points = 256 B, eight K-steps cycle
    timed(name, 0, idesc(0, M, N, 1, 1), sdesc(A_OFF, 128, 2048), 256 >> 4, 8, sdesc(B_OFF, 128, 2048), 256 >> 4, 8, 128, n_acc)


if __name__ == '__main__':
    kmajor('K-major SS M=128 N=32, 2 acc, lane 0 issues (as the engine does)', 32, 2)
    kmajor('K-major SS M=128 N=32, 2 acc, warp-uniform issue loop + elect.sync', 32, 2, uniform=1)
    kmajor('K-major SS M=128 N=32, 2 acc, TWO issuing warps', 32, 2, n_warps=2)
    kmajor('K-major SS M=128 N=32, 2 acc, FOUR issuing warps', 32, 2, n_warps=4)
    kmajor('K-major SS M=128 N=32, 2 acc, FOUR issuing warps, uniform', 32, 2, n_warps=4, uniform=1)
    kmajor('K-major SS M=128 N=8 (issue-rate floor)', 8, 2)
    kmajor('K-major SS M=128 N=64 K=16, one accumulator chain', 64, 1)
    kmajor('K-major SS M=128 N=64 K=16, 2 accumulators interleaved', 64, 2)
    kmajor('K-major SS M=128 N=64 K=16, 5 accumulators interleaved', 64, 5)
    kmajor('K-major SS M=128 N=16 K=16, 5 accumulators interleaved', 16, 5)
    kmajor('K-major SS M=128 N=128 K=16, 2 accumulators interleaved', 128, 2)
    kmajor('K-major SS M=128 N=256 K=16, one accumulator', 256, 1)
    kmajor('K-major SS M=64  N=64 K=16, 5 accumulators interleaved', 64, 5, M=64)
    kmajor('K-major TS (A from TMEM) M=128 N=64, one accumulator chain', 64, 1, ts=True)
    kmajor('K-major TS (A from TMEM) M=128 N=64, 5 accumulators interleaved', 64, 5, ts=True)
    kmajor('K-major TS (A from TMEM) M=128 N=128, 2 accumulators interleaved', 128, 2, ts=True)
    mnmajor('MN-major SS M=64  N=112 K=16 (weight gradient today), one accumulator', 64, 112)
    mnmajor('MN-major SS M=64  N=56  K=16, one accumulator', 64, 56)
    mnmajor('MN-major SS M=64  N=8   K=16 (bias gradient), one accumulator', 64, 8)
    mnmajor('MN-major SS M=128 N=112 K=16, one accumulator', 128, 112)
    mnmajor('MN-major SS M=128 N=112 K=16, 2 accumulators interleaved', 128, 112, 2)
    mnmajor('MN-major SS M=128 N=56  K=16, one accumulator', 128, 56)
    mnmajor('MN-major SS M=128 N=64  K=16, 2 accumulators interleaved', 128, 64, 2)
