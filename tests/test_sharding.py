"""Host-side tests of the sharded-image container (igneous_b200.sharding / shards) and of the
shard assembly of ImageShardDownsampleTask with the oracle standing in for the GPU kernels
(the GPU version of the same test is tests/test_tasks_gpu.py::test_image_shard_downsample_*)."""
import gzip
import struct

import numpy as np
import pytest

from igneous_b200 import sharding, shards


def test_compressed_morton_code_kats():
  # bits are dealt x, y, z from the LSB; a dimension drops out once its grid is covered
  assert sharding.compressed_morton_code((0, 0, 0), (4, 4, 4)) == 0
  assert sharding.compressed_morton_code((1, 2, 3), (4, 4, 4)) == 0b110101
  assert sharding.compressed_morton_code((3, 3, 3), (4, 4, 4)) == 63
  assert sharding.compressed_morton_code((2, 0, 0), (4, 1, 8)) == 0b100      # x0, z0, x1
  assert sharding.compressed_morton_code((3, 0, 7), (4, 1, 8)) == 0b11111
  assert sharding.compressed_morton_code((0, 0, 4), (4, 1, 8)) == 0b10000    # z2 comes after both x bits
  pts = np.array([[x, y, z] for z in range(3) for y in range(5) for x in range(2)])
  codes = sharding.compressed_morton_code(pts, (2, 5, 3))
  assert len(set(int(c) for c in codes)) == len(pts) and int(codes.max()) < 2 ** (1 + 3 + 2)
  with pytest.raises(ValueError):
    sharding.compressed_morton_code((4, 0, 0), (4, 4, 4))


def test_shard_file_layout_and_round_trip():
  spec = sharding.ShardingSpecification({"@type": sharding.SHARDING_TYPE, "preshift_bits": 1, "hash": "identity",
                                         "minishard_bits": 2, "shard_bits": 3,
                                         "minishard_index_encoding": "raw", "data_encoding": "raw"})
  # ids 40..47: >>1 = 20..23 -> minishards 0..3, shard (20>>2)&7 = 5
  chunks = {i: bytes([i]) * (i - 38) for i in range(40, 48)}
  assert {spec.locate(i) for i in chunks} == {(5, m) for m in range(4)}
  assert spec.shard_filename(5) == "5.shard"
  blob = spec.synthesize_shard(chunks)
  n = 16 * 4
  index = struct.unpack("<8Q", blob[:n])
  payload = sum(len(v) for v in chunks.values())
  assert index[0] == payload and index[-1] == len(blob) - n      # minishard indices follow the data
  first = np.frombuffer(blob[n + index[0]:n + index[1]], dtype="<u8").reshape(3, -1)
  assert first.tolist() == [[40, 1], [0, 0], [2, 3]]             # delta ids, gap-coded starts, sizes
  assert spec.chunk_ids(blob) == sorted(chunks)
  for cid, data in chunks.items():
    assert spec.read_chunk(blob, cid) == data
  assert spec.read_chunk(blob, 7) is None
  with pytest.raises(ValueError):
    spec.synthesize_shard({40: b"a", 0: b"b"})                    # two different shards
  gz = sharding.ShardingSpecification(dict(spec.to_dict(), minishard_index_encoding="gzip", data_encoding="gzip"))
  blob2 = gz.synthesize_shard(chunks)
  assert all(gz.read_chunk(blob2, cid) == data for cid, data in chunks.items())
  s, e = struct.unpack_from("<QQ", blob2, 0)
  assert np.frombuffer(gzip.decompress(blob2[n + s:n + e]), dtype="<u8").reshape(3, -1)[0].tolist() == [40, 1]
  with pytest.raises(NotImplementedError):
    sharding.ShardingSpecification(dict(spec.to_dict(), hash="murmurhash3_x86_128"))


def test_image_shard_shape_from_spec():
  f = shards.image_shard_shape_from_spec
  assert list(f({"preshift_bits": 4, "minishard_bits": 0}, (512, 512, 512), (64, 64, 64))) == [256, 128, 128]
  assert list(f({"preshift_bits": 2, "minishard_bits": 1}, (512, 512, 512), (64, 64, 64))) == [128, 128, 128]
  assert list(f({"preshift_bits": 9, "minishard_bits": 3}, (512, 512, 64), (64, 64, 64))) == [512, 512, 64]
  assert list(f({"preshift_bits": 3, "minishard_bits": 0}, (100, 512, 512), (64, 64, 64))) == [128, 128, 128]
  assert list(f({"preshift_bits": 5, "minishard_bits": 0}, (100, 512, 512), (64, 64, 64))) == [128, 256, 256]
  with pytest.raises(ValueError):
    f({"preshift_bits": 60, "minishard_bits": 4}, (512, 512, 512), (64, 64, 64))


@pytest.mark.parametrize("size,chunk,target", [((512, 512, 512), (64, 64, 64), 64 ** 3 * 4 * 16),
                                               ((2048, 2048, 256), (128, 128, 16), int(3.5e9)),
                                               ((100, 60, 33), (32, 32, 32), 1 << 20)])
def test_create_sharded_image_info_invariants(size, chunk, target):
  spec = sharding.create_sharded_image_info(size, chunk, "raw", np.uint32, uncompressed_shard_bytesize=target)
  _, nb = sharding.grid_bits(size, chunk)
  assert spec["preshift_bits"] + spec["minishard_bits"] + spec["shard_bits"] == sum(nb)
  assert (16 << spec["minishard_bits"]) <= 8192 or spec["preshift_bits"] + spec["minishard_bits"] == sum(nb)
  shard = shards.image_shard_shape_from_spec(spec, size, chunk)
  assert int(np.prod([int(v) for v in shard])) * 4 <= max(target, int(np.prod(chunk)) * 4)
  # every chunk of one shard box hashes to one shard file, and boxes do not share files
  s = sharding.ShardingSpecification(spec)
  grid, _ = sharding.grid_bits(size, chunk)
  per = [int(a) // int(c) for a, c in zip(shard, chunk)]
  owner = {}
  for z in range(grid[2]):
    for y in range(grid[1]):
      for x in range(grid[0]):
        box = (x // per[0], y // per[1], z // per[2])
        no = s.locate(sharding.compressed_morton_code((x, y, z), grid))[0]
        assert owner.setdefault(no, box) == box


@pytest.mark.parametrize("factor,shape,offset,chunk,target", [
  ((2, 2, 1), (256, 192, 96), (0, 0, 0), (32, 32, 32), 32 ** 3 * 4 * 8),
  ((2, 2, 2), (128, 128, 128), (3, 5, 7), (32, 32, 32), 32 ** 3 * 4 * 8),
  ((2, 2, 1), (128, 128, 160), (0, 0, 0), (32, 32, 32), 32 ** 3 * 4 * 4),     # coarse shards taller than a task
])
def test_image_shard_downsample_assembly_with_oracle_kernels(oracle, tmp_path, monkeypatch, factor, shape,
                                                             offset, chunk, target):
  """Shard assembly only: pooling / renumber / remap are replaced by the oracle and numpy."""
  from igneous_b200 import tinybrain, fastremap
  import igneous_b200.task_creation as tc
  from igneous_b200._compat import CloudVolume, LocalTaskQueue

  def renumber(a, preserve_zero=True, in_place=False):
    u, inv = np.unique(a, return_inverse=True)
    new = np.arange(len(u), dtype=np.uint32) + (0 if u[0] == 0 else 1)
    return np.asfortranarray(new[inv].reshape(a.shape).astype(np.uint16)), {int(o): int(n) for o, n in zip(u, new)}

  def remap(a, table, preserve_missing_labels=False):
    lut = np.zeros(max(table) + 1, dtype=np.uint64)
    for k, v in table.items():
      lut[k] = v
    return lut[a].astype(a.dtype)

  monkeypatch.setattr(tinybrain, "downsample_segmentation",
                      lambda img, f, num_mips=1, sparse=False: oracle.downsample_segmentation(img, f, num_mips=num_mips, sparse=sparse))
  monkeypatch.setattr(fastremap, "renumber", renumber)
  monkeypatch.setattr(fastremap, "remap", remap)
  seg = oracle.synth_seg(shape, pitch=24).astype(np.uint32)[..., np.newaxis]
  path = "file://" + str(tmp_path / "seg")
  CloudVolume.from_numpy(seg, vol_path=path, resolution=(16, 16, 40), voxel_offset=offset, chunk_size=chunk,
                         layer_type="segmentation")
  tasks = tc.create_image_shard_downsample_tasks(path, mip=0, num_mips=2, factor=factor, memory_target=target)
  LocalTaskQueue(parallel=1).insert_all(tasks)
  cv = CloudVolume(path)
  assert cv.available_mips == [0, 1, 2]
  assert all("sharding" in cv.scales[m] for m in (1, 2)) and "sharding" not in cv.scales[0]
  want = oracle.downsample_segmentation(seg, tuple(factor) + (1,), num_mips=2)
  for m in (1, 2):
    cv.mip = m
    assert np.array_equal(cv[cv.meta.bounds(m)], want[m - 1]), m
  assert cv.provenance.processing[-1]["method"]["task"] == "ImageShardDownsampleTask"
  with pytest.raises(NotImplementedError):
    cv[cv.meta.bounds(2)] = want[1]
