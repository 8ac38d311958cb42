"""Regenerates tests/golden/*.npz from the CPU oracle.

The reference's own kernels (tinybrain / cc3d / zmesh wheels) cannot be imported in
this image, so these vectors freeze the ORACLE's outputs on small seeded inputs; the
oracle itself is pinned by the reference's known answers in tests/test_oracle.py.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
  seg = O.synth_seg((48, 40, 24), pitch=16, num_ids=7, seed=1)
  img = O.synth_image((48, 40, 6), seed=2)
  mode = O.downsample_segmentation(seg, (2, 2, 1), num_mips=3)
  avg = O.downsample_with_averaging(img, (2, 2, 1), num_mips=5)
  cc, n = O.connected_components(seg, return_N=True)
  dust = O.dust(seg, 40)
  ren, mapping = O.renumber(seg)
  tl, tv = O.marching_cubes(seg[:25, :21, :13])
  W = O.WeldedMeshes(tl, tv)
  lab = W.ids()[0]
  v, f = W.get(lab, (16, 16, 40), True)
  simp, rounds = O.simplify_welded(W, (16, 16, 40), 4, 1e9, True)
  np.savez_compressed(
    os.path.join(HERE, "hotpath_small.npz"), seg=seg, img=img, mode1=mode[0], mode2=mode[1], mode3=mode[2],
    avg1=avg[0], avg2=avg[1], avg3=avg[2], avg4=avg[3], avg5=avg[4], cc=cc.astype(np.uint32), n=np.uint64(n),
    dust40=dust, renumber=ren, mesh_label=np.uint64(lab), mesh_vertices=v, mesh_faces=f,
    simp_vertices=simp[lab][0], simp_faces=simp[lab][1], n_triangles=np.uint64(len(tl)))
  print("wrote", os.path.join(HERE, "hotpath_small.npz"))
  # block pooling (factors 1 or 2 per axis other than (2,2,1)), SURVEY 8(f) row 3
  blocks = {}
  for name, arrs in (("mode222", O.downsample_segmentation(seg, (2, 2, 2), num_mips=2)),
                     ("smode222", O.downsample_segmentation(seg, (2, 2, 2), num_mips=2, sparse=True)),
                     ("mode122", O.downsample_segmentation(seg, (1, 2, 2), num_mips=2)),
                     ("avg222", O.downsample_with_averaging(img, (2, 2, 2), num_mips=2)),
                     ("savg221", O.downsample_with_averaging(np.where(img > 128, img, 0).astype(np.uint8), (2, 2, 1),
                                                            num_mips=2, sparse=True))):
    for k, a in enumerate(arrs):
      blocks["%s_%d" % (name, k + 1)] = a
  np.savez_compressed(os.path.join(HERE, "pooling_blocks.npz"), **blocks)
  print("wrote", os.path.join(HERE, "pooling_blocks.npz"))


if __name__ == "__main__":
  main()
