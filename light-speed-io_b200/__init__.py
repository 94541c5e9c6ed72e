"""lsio-b200: B200-native drop-in for light-speed-io's batched byte-range read."""
from . import _native  # noqa: F401  (raises if liblsio_b200.so is missing: no CPU fallback)
from . import cuda_ops  # noqa: F401
from . import codecs  # noqa: F401
from . import zarr_shard  # noqa: F401
from . import zarr_v3  # noqa: F401
from .aligned_bytes import AlignedBytes, AlignedBytesError, AlignedBytesMut  # noqa: F401
from .io_uring_cuda import Chunk, Disconnected, IoError, IoUringCuda, Receiver, RecvTimeoutError  # noqa: F401
