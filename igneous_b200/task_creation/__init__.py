from .common import FinelyDividedTaskIterator, get_bounds, num_tasks, operator_contact
from .image import (create_downsampling_tasks, create_image_shard_downsample_tasks, num_mips_from_memory_target, create_ccl_face_tasks,
                    create_ccl_equivalence_tasks, create_ccl_relabel_tasks, MEMORY_TARGET)
from .mesh import create_meshing_tasks
