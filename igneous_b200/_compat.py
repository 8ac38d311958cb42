"""Storage / queue layer used by the task mirror: the real cloud-volume,
cloud-files and task-queue packages when importable, otherwise the minimal
file:// stand-ins of igneous_b200.storage (never both)."""
try:  # pragma: no cover - not installable in the build image
  from cloudvolume import CloudVolume, EmptyVolumeException
  from cloudvolume.exceptions import InfoUnavailableError
  from cloudvolume.lib import Vec, Bbox, min2, max2
  from cloudfiles import CloudFiles
  from taskqueue import queueable, RegisteredTask, LocalTaskQueue
  USING_STANDINS = False
except ImportError:
  from .storage import (CloudVolume, EmptyVolumeException, InfoUnavailableError, Vec, Bbox, min2, max2,
                        CloudFiles, queueable, RegisteredTask, LocalTaskQueue)
  USING_STANDINS = True
