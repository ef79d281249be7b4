"""How many reduction lane-ops of the forward kernel hit the slot of a neighbouring lane?  CPU replay (build container or
any host): the oracle's chain positions of a synthetic window at the headline event density, tile-sorted like
csrc/tef_cm_sort.cu does, cut into warps of 32, keyed like splat_inside_1hot_warp keys its top-row reduction.

    python tests/tools/merge_replay.py [uniform|edges]

Prints the fraction of lanes whose key equals the previous lane's (what one round of merge_equal_neighbours can use) and the
number of distinct keys per lane (what a full per-warp match could reach).  DESIGN.md decision 13 quotes these numbers."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
from oracle import cm_oracle as orc  # noqa: E402  (a measurement script, not product code)


def main():
    dist = sys.argv[1] if len(sys.argv) > 1 else "uniform"
    H, W, P = 240, 320, 10
    N = int(3.255 * H * W)                                   # 1 M events on 480x640 = 3.255 events per pixel and window
    wl = dict(B=1, P=P, N=N, Nd=0, H=H, W=W, F=1, S=1, mode="two", sigma=3.0, dist=dist, warping="Iterative")
    seq = bench.fast_sequence(3, wl)
    o = orc.iterative(orc.make_cfg(1, H, W, P, 1), seq["flows"], seq["events"], seq["masks"], seq["d_events"], seq["d_masks"],
                      np.float32, want_grad=False, want_iwe=False, want_nodes=True)
    nodes, alive = o["nodes"][0], o["alive"][0]              # [P+1, E, 2] (y, x), [P+1, E]
    lanes = adjacent = distinct = 0
    for t in range(P):
        ev = seq["events"][t][0].numpy()
        n = ev.shape[0]
        y0, x0, pol = ev[:, 1].astype(int), ev[:, 2].astype(int), (ev[:, 3] < 0).astype(int)
        key = ((y0 >> 3) * ((W + 15) // 16) + (x0 >> 4)) * 128 + ((y0 & 7) << 4) + (x0 & 15)       # sort_bin()
        order = np.argsort(key, kind="stable")
        idx = np.arange(t * n, (t + 1) * n)[order]
        ok = alive[:, idx].all(0)                            # border compensation: alive at every reference time
        for tr in range(max(0, t - 4), min(P, t + 5) + 1):   # mode two: the reference times this window feeds
            pos = nodes[tr, idx]
            y, x = np.floor(pos[:, 0]).astype(np.int64), np.floor(pos[:, 1]).astype(np.int64)
            phase = x & 1
            slot = (((phase * 2 + pol[order]) * H + y) * (W + 4) + x + phase) >> 1
            slot = np.where(ok, slot, -1 - np.arange(n))     # lanes without work never merge
            nw = n // 32
            k, live = slot[:nw * 32].reshape(nw, 32), ok[:nw * 32].reshape(nw, 32)
            adjacent += int(((k[:, 1:] == k[:, :-1]) & live[:, 1:]).sum())
            lanes += int(live.sum())
            s = np.sort(k, axis=1)
            distinct += int(((s[:, 1:] != s[:, :-1]).sum(1) + 1 - (~live).sum(1)).clip(min=0).sum())
    print("%s events: %d live lanes, %.1f %% repeat the previous lane's slot, %.2f distinct slots per lane"
          % (dist, lanes, 100.0 * adjacent / lanes, distinct / lanes))


if __name__ == "__main__":
    main()
