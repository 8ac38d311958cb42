"""Task-level tests on a file:// Precomputed layer, mirroring the reference's
own integration tests (test/test_tasks.py, test/test_ccl_tasks.py) with the
GPU libraries behind the tasks and the oracle as the checker."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _layer(tmp_path, data, layer_type, chunk=(64, 64, 64), res=(1, 1, 1), offset=(0, 0, 0), name="layer"):
  from igneous_b200._compat import CloudVolume
  path = "file://" + str(tmp_path / name)
  CloudVolume.from_numpy(data, vol_path=path, resolution=res, voxel_offset=offset, chunk_size=chunk,
                         layer_type=layer_type, max_mip=0)
  return path


def test_downsample_task_config_c1(ctx, oracle, tmp_path):
  """BASELINE config C1: DownsampleTask mip0->mip1 (2x2x1 mode) on 128x128x64 uint32,
  LocalTaskQueue(parallel=1), file:// layer."""
  import igneous_b200.task_creation as tc
  from igneous_b200._compat import CloudVolume, LocalTaskQueue
  seg = oracle.synth_seg((128, 128, 64), pitch=16, num_ids=64)[..., np.newaxis]
  path = _layer(tmp_path, seg, "segmentation")
  tq = LocalTaskQueue(parallel=1)
  tq.insert_all(tc.create_downsampling_tasks(path, mip=0, num_mips=1, compress="gzip"))
  cv = CloudVolume(path)
  assert len(cv.available_mips) == 2
  assert list(cv.meta.volume_size(1)) == [64, 64, 64]
  want, = oracle.downsample_segmentation(seg, (2, 2, 1, 1), num_mips=1)
  cv.mip = 1
  assert np.array_equal(cv[cv.meta.bounds(1)], want)
  assert cv.provenance.processing[-1]["method"]["task"] == "DownsampleTask"


def test_downsample_task_volumetric_factor(ctx, oracle, tmp_path):
  """--volumetric: factor (2,2,2) mode pooling through DownsampleTask."""
  import igneous_b200.task_creation as tc
  from igneous_b200._compat import CloudVolume, LocalTaskQueue
  seg = oracle.synth_seg((256, 256, 256), pitch=16, num_ids=64)[..., np.newaxis]
  path = _layer(tmp_path, seg, "segmentation")
  LocalTaskQueue(parallel=1).insert_all(tc.create_downsampling_tasks(path, mip=0, num_mips=2, factor=(2, 2, 2)))
  cv = CloudVolume(path)
  assert [list(map(int, cv.meta.volume_size(m))) for m in cv.available_mips] == \
      [[256, 256, 256], [128, 128, 128], [64, 64, 64]]
  want = oracle.downsample_segmentation(seg, (2, 2, 2, 1), num_mips=2)
  for m in (1, 2):
    cv.mip = m
    assert np.array_equal(cv[cv.meta.bounds(m)], want[m - 1])


@pytest.mark.parametrize("compress", [None, "gzip", "br"])
def test_downsample_no_offset_average_pyramid(ctx, oracle, tmp_path, compress):
  """test/test_tasks.py:29-71: 4 average mips of a 1024x1024x128 uint8 image."""
  import igneous_b200.task_creation as tc
  from igneous_b200._compat import CloudVolume, LocalTaskQueue
  img = oracle.synth_image((1024, 1024, 128))[..., np.newaxis]
  path = _layer(tmp_path, img, "image")
  LocalTaskQueue().insert_all(tc.create_downsampling_tasks(path, mip=0, num_mips=4, compress=compress))
  cv = CloudVolume(path)
  assert len(cv.available_mips) == 5
  sizes = [list(cv.meta.volume_size(m)) for m in range(5)]
  assert sizes == [[1024, 1024, 128], [512, 512, 128], [256, 256, 128], [128, 128, 128], [64, 64, 128]]
  want = oracle.downsample_with_averaging(img, (2, 2, 1, 1), num_mips=4)
  for m in range(1, 5):
    cv.mip = m
    assert np.array_equal(cv[cv.meta.bounds(m)], want[m - 1])


def test_downsample_with_offset_and_missing_chunk(ctx, oracle, tmp_path):
  """test/test_tasks.py:249-322: voxel offset; fill_missing turns holes into zeros."""
  import igneous_b200.task_creation as tc
  from igneous_b200._compat import CloudVolume, CloudFiles, LocalTaskQueue, EmptyVolumeException
  img = oracle.synth_image((512, 512, 128))[..., np.newaxis]
  path = _layer(tmp_path, img, "image", offset=(3, 7, 11))
  cf = CloudFiles(path)
  cf.delete(["1_1_1/67-131_7-71_11-75"])
  with pytest.raises(EmptyVolumeException):
    LocalTaskQueue().insert_all(tc.create_downsampling_tasks(path, mip=0, num_mips=3))
  LocalTaskQueue().insert_all(tc.create_downsampling_tasks(path, mip=0, num_mips=3, fill_missing=True))
  holed = img.copy()
  holed[64:128, 0:64, 0:64] = 0
  want = oracle.downsample_with_averaging(holed, (2, 2, 1, 1), num_mips=3)
  cv = CloudVolume(path)
  for m in range(1, 4):
    cv.mip = m
    assert np.array_equal(cv[cv.meta.bounds(m)], want[m - 1])


def _checker():
  data = np.zeros((512, 512, 128), dtype=np.uint8)
  i = 1
  for x in range(8):
    for y in range(8):
      for z in range(2):
        data[64 * x:64 * (x + 1), 64 * y:64 * (y + 1), 64 * z:64 * (z + 1)] = i
        i += 1
  return data


def _run_ccl(path, dest, shape, **kw):
  import igneous_b200.task_creation as tc
  from igneous_b200 import tasks
  from igneous_b200._compat import LocalTaskQueue
  tq = LocalTaskQueue()
  tq.insert_all(tc.create_ccl_face_tasks(path, mip=0, shape=shape, **kw))
  tq.insert_all(tc.create_ccl_equivalence_tasks(path, mip=0, shape=shape, **kw))
  tasks.create_relabeling(path, mip=0, shape=shape)
  tq.insert_all(tc.create_ccl_relabel_tasks(path, dest, mip=0, shape=shape, **kw))


@pytest.mark.parametrize("upper", (None, 255, 0))
@pytest.mark.parametrize("dust_threshold", [0, 64 ** 3 + 1])
def test_ccl_tasks_checker(ctx, tmp_path, upper, dust_threshold):
  """test/test_ccl_tasks.py:111-211: file inventory and label known answers."""
  from igneous_b200._compat import CloudVolume, CloudFiles
  path = _layer(tmp_path, _checker(), "image", chunk=(128, 128, 64), name="src")
  dest = "file://" + str(tmp_path / "dest")
  _run_ccl(path, dest, (128, 128, 128), threshold_lte=upper, dust_threshold=dust_threshold)
  cf = CloudFiles(path)
  faces = cf.list("1_1_1/ccl/faces")
  want_faces = sorted("1_1_1/ccl/faces/%d-%d-0-%s.npy" % (x, y, k) for x in range(4) for y in range(4)
                      for k in ("xy", "xz", "yz"))
  assert sorted(faces) == want_faces
  assert sorted(cf.list("1_1_1/ccl/equivalences")) == sorted(
    "1_1_1/ccl/equivalences/%d-%d-0.json" % (x, y) for x in range(4) for y in range(4))
  cc = CloudVolume(dest)[:][:, :, :, 0]
  uniq = np.unique(cc)
  if dust_threshold > 0:
    assert list(uniq) == ([0] if upper in (None, 0) else [1])
  elif upper is None:
    assert np.array_equal(uniq, np.arange(1, 129))
  elif upper == 255:
    assert list(uniq) == [1]
  else:
    assert list(uniq) == [0]


def test_ccl_tasks_equal_single_shot_ccl(ctx, oracle, tmp_path):
  """test/test_ccl_tasks.py:213-249 on a synthetic connectomics-like volume: the
  4-pass out-of-core result equals one whole-volume CCL after canonical renumbering."""
  from igneous_b200._compat import CloudVolume
  seg = oracle.synth_seg((256, 192, 128), pitch=32, num_ids=6)
  path = _layer(tmp_path, seg[..., np.newaxis], "segmentation", chunk=(128, 64, 64), name="src")
  dest = "file://" + str(tmp_path / "dest")
  _run_ccl(path, dest, (128, 128, 128))
  got = CloudVolume(dest)[:][:, :, :, 0]
  want = oracle.connected_components(seg)
  a, _ = oracle.renumber(got)
  b, _ = oracle.renumber(want)
  assert np.array_equal(a, b)


@pytest.mark.parametrize("compress", ("gzip", "br"))
def test_mesh_task_box(ctx, tmp_path, compress):
  """test/test_tasks.py:407-431: 62^3 box, remap_table, fragment file name."""
  from igneous_b200.tasks import MeshTask
  from igneous_b200._compat import CloudVolume, CloudFiles
  from igneous_b200 import zmesh
  data = np.zeros((64, 64, 64, 1), dtype=np.uint32)
  data[1:-1, 1:-1, 1:-1, :] = 1
  path = _layer(tmp_path, data, "segmentation")
  cv = CloudVolume(path)
  cv.info["mesh"] = "mesh"
  cv.commit_info()
  MeshTask(shape=(64, 64, 64), offset=(0, 0, 0), layer_path=path, mip=0, remap_table={"1": "10"},
           low_padding=0, high_padding=1, compress=compress).execute()
  cf = CloudFiles(path)
  assert cf.list("mesh/") == ["mesh/10:0:0-64_0-64_0-64"]
  frag = zmesh.Mesh.from_precomputed(cf.get("mesh/10:0:0-64_0-64_0-64"))
  assert 0 < len(frag.faces) <= 46124 // 100 + 2
  # closed (dataset edges are zero padded) and positioned in dataset coordinates
  f = frag.faces.astype(np.int64)
  e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
  assert np.array_equal(np.sort(e[:, 0] * (1 << 32) + e[:, 1]), np.sort(e[:, 1] * (1 << 32) + e[:, 0]))
  assert frag.vertices.min() >= 0.99 and frag.vertices.max() <= 63.01


def test_mesh_task_object_ids_and_creator(ctx, oracle, tmp_path):
  """test/test_tasks.py:433-462 + create_meshing_tasks: ids filters, mesh info, spatial index."""
  import igneous_b200.task_creation as tc
  from igneous_b200.tasks import MeshTask
  from igneous_b200._compat import CloudVolume, CloudFiles, LocalTaskQueue
  data = np.zeros((64, 64, 64, 1), dtype=np.uint32)
  data[1:-1, 1:-1, 1:-1, :] = 1
  data[1:-1, 1:-1, 32:63, :] = 2
  path = _layer(tmp_path, data, "segmentation")
  cv = CloudVolume(path)
  cv.info["mesh"] = "mesh"
  cv.commit_info()
  MeshTask(shape=(64, 64, 64), offset=(0, 0, 0), layer_path=path, mip=0, exclude_object_ids=[2]).execute()
  cf = CloudFiles(path)
  assert cf.list("mesh/") == ["mesh/1:0:0-64_0-64_0-64"]
  MeshTask(shape=(64, 64, 64), offset=(0, 0, 0), layer_path=path, mip=0, object_ids=[2]).execute()
  assert cf.get("mesh/2:0:0-64_0-64_0-64") is not None
  # unsimplified fragments equal the oracle's marching cubes of the padded volume
  tasks = tc.create_meshing_tasks(path, mip=0, shape=(32, 64, 64), simplification=False, mesh_dir="m2")
  assert len(tasks) == 2
  LocalTaskQueue().insert_all(tasks)
  info = cf.get_json("m2/info")
  assert info["@type"] == "neuroglancer_legacy_mesh" and info["chunk_size"] == [32, 64, 64]
  names = cf.list("m2/")
  assert "m2/1:0:0-32_0-64_0-64" in names and "m2/2:0:32-64_0-64_0-64" in names
  assert any(n.endswith(".spatial") for n in names)


def test_downsample_task_compressed_segmentation_encoding(ctx, oracle, tmp_path):
  """a DownsampleTask whose new mips are `compressed_segmentation` layers
  (create_downsampling_tasks(encoding=...), igneous/task_creation/image.py:284-286): the chunks
  are written by the device codec, byte-identical to the oracle encoder, and read back through
  the device decoder."""
  import gzip
  import igneous_b200.task_creation as tc
  from igneous_b200._compat import CloudVolume, CloudFiles, LocalTaskQueue
  seg = oracle.synth_seg((256, 256, 64), pitch=16, num_ids=64)[..., np.newaxis]
  path = _layer(tmp_path, seg, "segmentation")
  LocalTaskQueue(parallel=1).insert_all(tc.create_downsampling_tasks(
    path, mip=0, num_mips=2, encoding="compressed_segmentation", compress="gzip"))
  cv = CloudVolume(path)
  assert [s["encoding"] for s in cv.info["scales"]] == ["raw", "compressed_segmentation", "compressed_segmentation"]
  want = oracle.downsample_segmentation(seg, (2, 2, 1, 1), num_mips=2)
  for m in (1, 2):
    cv.mip = m
    assert np.array_equal(cv[cv.meta.bounds(m)], want[m - 1])
  # the stored chunk IS the oracle's stream
  cf = CloudFiles(path)
  name = [n for n in cf.list(cv.info["scales"][1]["key"])][0]
  raw = cf.get(name)
  box = [tuple(int(v) for v in part.split("-")) for part in name.split("/")[-1].split("_")]
  chunk = want[0][box[0][0]:box[0][1], box[1][0]:box[1][1], box[2][0]:box[2][1]]
  assert raw == oracle.cseg_encode(np.asfortranarray(chunk)).tobytes()


@pytest.mark.parametrize("factor,shape,offset,encoding", [
  ((2, 2, 1), (256, 192, 96), (0, 0, 0), "raw"),
  ((2, 2, 2), (128, 128, 128), (3, 5, 7), "raw"),                       # test_downsample_with_offset_sharded_2x2x2
  ((2, 2, 1), (128, 128, 160), (0, 0, 0), "compressed_segmentation"),   # coarse shards taller than a task
])
def test_image_shard_downsample_task(ctx, oracle, tmp_path, factor, shape, offset, encoding):
  """create_image_shard_downsample_tasks + ImageShardDownsampleTask
  (test/test_tasks.py:73-124,154-245): sharded mips equal the pooling of the whole volume;
  renumber -> pooling -> remap and the chunk codec run on the GPU, the shard container on the host."""
  import igneous_b200.task_creation as tc
  from igneous_b200._compat import CloudVolume, LocalTaskQueue
  seg = oracle.synth_seg(shape, pitch=24, num_ids=1 << 20).astype(np.uint32)[..., np.newaxis]
  seg[seg != 0] += np.uint32(1 << 24)            # ids that do not fit the renumbered dtype
  path = "file://" + str(tmp_path / "seg")
  CloudVolume.from_numpy(seg, vol_path=path, resolution=(16, 16, 40), voxel_offset=offset, chunk_size=(32, 32, 32),
                         layer_type="segmentation")
  tasks = tc.create_image_shard_downsample_tasks(path, mip=0, num_mips=2, factor=factor, encoding=encoding,
                                                 memory_target=32 ** 3 * 4 * (4 if shape[2] == 160 else 8))
  LocalTaskQueue(parallel=1).insert_all(tasks)
  cv = CloudVolume(path)
  assert cv.available_mips == [0, 1, 2]
  assert all(cv.scales[m]["sharding"]["@type"] == "neuroglancer_uint64_sharded_v1" for m in (1, 2))
  assert all(cv.scales[m]["encoding"] == encoding for m in (1, 2))
  want = oracle.downsample_segmentation(seg, tuple(factor) + (1,), num_mips=2)
  for m in (1, 2):
    cv.mip = m
    assert np.array_equal(cv[cv.meta.bounds(m)], want[m - 1]), m
