"""Arithmetic behind DESIGN.md sections 4.1, 4.2, 4.6 and 5: on-chip and HBM floors of the HEX8 elasticity assembly on B200.

    python tools/roofline_model.py [N]        # N^3 cells, default 100 (cfg 2)

Machine numbers: 148 SMs at 1.965 GHz; FP64 64 FMA/clk/SM (tools/microbench.cu: 36.8 TFLOP/s DFMA, 37.1 TFLOP/s DMMA, one
shared pipe); shared memory 128 B/clk/SM = one wavefront per clock; HBM 6550 GB/s (MEASURED_PEAKS.json)."""
import sys

SM, GHZ, FMA_CLK, HBM = 148, 1.965, 64, 6550e9


def ms(cycles_per_sm):
    return cycles_per_sm / (GHZ * 1e9) * 1e3


def main(N=100):
    C, nodes = N ** 3, (N + 1) ** 3
    nnz = 9 * (3 * (N + 1) - 2) ** 3
    n = 3 * nodes
    b_asm = 8 * nnz + 8 * n + 4 * 8 * C + (8 * 3 + 8 * 3) * nodes
    print(f"cfg: {N}^3 cells, {n} DOF, nnz {nnz}; algorithmic bytes {b_asm / 1e9:.3f} GB -> {b_asm / HBM * 1e3:.3f} ms at 100 %, "
          f"{b_asm / HBM / 0.6 * 1e3:.3f} ms at the 60 % target")
    # FP64 work per cell (FMA): phase 1 ~ 8 q x 236, products 64 pairs x 9 x 8 q, conversion ~ 20 per block, residual 8 x 8 x 9
    p1, prod, conv, resid = 8 * 236, 64 * 9 * 8, 64 * 20, 8 * 8 * 9
    sym = 36 * 9 * 8
    for name, fma in (("all 64 blocks", p1 + prod + conv + resid), ("symmetric 36 blocks", p1 + sym + conv + resid),
                      ("DMMA tiles (24 x 256) + phase 1", p1 + 24 * 256)):
        print(f"FP64 floor, {name:32s}: {fma:6d} FMA/cell -> {ms(C * fma / FMA_CLK / SM):.3f} ms at 100 % of the pipe")
    # shared-memory wavefronts per cell (one per 128 B): accumulate = read-modify-write of 64 x 9 doubles
    acc = 64 * 9 * 8 * 2 / 128
    print(f"shared-memory floor of a read-modify-write accumulator: {acc:.0f} wavefronts/cell -> {ms(C * acc / SM):.3f} ms; "
          f"with the element kernel's phase 1 + fragments (64): {ms(C * (acc + 64) / SM):.3f} ms")
    # two-kernel path: bytes through HBM (ncu, profiles/r01_traffic.json at N = 100 scales with C)
    staged = (4.743e9 + 0.115e9 + 5.320e9 + 1.937e9) * C / 1e6
    print(f"two-kernel path: {staged / 1e9:.2f} GB through HBM -> {staged / HBM * 1e3:.3f} ms floor (measured 2.46 ms at N = 100)")
    # owner-computes fusion with 64-node patches: 125 cells per 64 owned nodes
    red = 125 / 64
    print(f"owner-computes fusion: phase 1 x {red:.2f}; CUDA-core FP64 {ms(C * (p1 * red + prod + conv + resid) / FMA_CLK / SM):.3f} ms, "
          f"tensor-core variant {ms(C * red * (p1 + 24 * 256) / FMA_CLK / SM):.3f} ms at 100 % (measured 3.92 / 4.64 ms)")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 100)
