"""One tiny train step + one style-transfer call on the cluster kernels, for compute-sanitizer (memcheck / racecheck / synccheck / initcheck).

    compute-sanitizer --tool memcheck python scripts/sanitize_case.py 512 6 72

argv: H T B [precision [cell_type]].  cell_type GRU with H = 256: gru_cluster_fwd_kernel / gru_cluster_bwd_kernel (4-CTA clusters).  72 rows = one full 64-row group + a ragged 8-row group; H = 256 -> 8-CTA clusters, H = 512 -> 16-CTA clusters
(rec_cluster_fwd2_kernel, rec_cluster_bwd_kernel / rec_cluster_bwd4_kernel).  Prints the loss so that a sanitizer run can be compared with a plain one.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from midi_vae_b200 import Engine  # noqa: E402
from tests import util            # noqa: E402  (shared batch / weight builders; nothing of the oracle is executed)


def main():
    H, T, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    precision = sys.argv[4] if len(sys.argv) > 4 else "bf16"
    cell = sys.argv[5] if len(sys.argv) > 5 else "LSTM"
    ecfg, _ = util.make_cfgs(T=T, H=H, L=32, feedback="teacher_forced", precision=precision, max_batch=B, cell_type=cell)
    eng = Engine(ecfg, 0)
    eng.set_weights(util.make_weights(ecfg))
    r, hist, eps, sw = util.make_batch(ecfg, B, weights=True)
    for _ in range(2):
        m = eng.train_on_batch(r.pitch, r.instr, r.velocity, r.style, hist, eps, sw)
    P, Ii, Vv = eng.style_transfer(r.pitch, r.instr, r.velocity, 0, 1, None, "as_wired")
    print(f"sanitize_case {cell} H={H} T={T} B={B} {precision}: loss {m['loss']:.6f}, launches {eng.launch_count()}, pitch checksum {int(P.sum())}")
    eng.close()


if __name__ == "__main__":
    main()
