"""CPU model of what the paper-flavour kernel streams on a named synthetic graph (no GPU needed).

    python profiles/stream_model.py [arxiv|squirrel|cora]

Prints the numbers DESIGN.md §4.2 quotes: streamed entries with the cheaper-side rule against the both-sides model of
SURVEY.md §8d, the distribution of streamed list lengths (most LISTS are short, most ELEMENTS sit in long lists — the
reason for the two streaming paths), the classes by degree of the tested endpoint, and how much of the stream belongs to
edges above the cooperative / split thresholds.
"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "discrete-curvature-rewiring_b200"))
from dcr.synth import csr_from_edge_index, named_graph  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "arxiv"
ei, n = named_graph(name)
rowptr, col = csr_from_edge_index(ei, n)
rowptr = rowptr.astype(np.int64)
deg = np.diff(rowptr)
S = np.add.reduceat(deg[col], rowptr[:-1]) * (deg > 0)          # S_v = sum of the neighbours' degrees
m = ei[0] < ei[1]
I, J = ei[0][m], ei[1][m]
di, dj = deg[I], deg[J]
ca, cb = S[J] - di, S[I] - dj                                    # 2-hop entries behind j / behind i
swapped = cb < ca
stream = np.where(swapped, cb, ca)
da = np.where(swapped, dj, di)
B = np.where(swapped, I, J)
triv = np.minimum(di, dj) <= 1
print(f"{name}: n={n} E={I.size} max degree {deg.max()}; trivial edges (deg_min <= 1): {triv.sum()}")
print(f"streamed entries, cheaper side only: {stream[~triv].sum() / 1e9:.3f} G; both sides (SURVEY §8d model): "
      f"{(ca + cb)[~triv].sum() / 1e9:.3f} G")
wb = np.bincount(B[~triv], minlength=n)
w = np.repeat(wb, deg)
dm = deg[col]
bins = [0, 8, 16, 32, 64, 128, 256, 1024, 1 << 30]
cnt, _ = np.histogram(dm, bins=bins, weights=w)
el, _ = np.histogram(dm, bins=bins, weights=w * dm)
print("list length bins  ", bins[:-1])
print("share of lists   %", np.round(100 * cnt / cnt.sum(), 1))
print("share of elements%", np.round(100 * el / el.sum(), 1))
for lo, hi, tag in ((0, 128, "d_a <= 128 (warp-private table)"), (128, 1 << 30, "d_a > 128 (group kernel)")):
    k = (~triv) & (da > lo) & (da <= hi)
    print(f"{tag}: {k.sum()} edges, {stream[k].sum() / 1e9:.3f} G entries")
for thr, tag in ((8192, "cooperative threshold for d_a <= 128"), (24576, "cooperative threshold"), (98304, "split threshold")):
    k = (~triv) & (stream > thr)
    print(f"stream > {thr} ({tag}): {k.sum()} edges ({100 * k.sum() / (~triv).sum():.2f} %), "
          f"{100 * stream[k].sum() / stream[~triv].sum():.1f} % of the stream; largest {stream.max()}")
