"""FileMapper — block hash -> file path, same layout as the reference
(kv_connectors/llmd_fs_backend/llmd_fs_backend/file_mapper.py:18-87) so existing caches stay valid."""
from __future__ import annotations


class FileMapper:
    def __init__(self, root_dir: str, model_name: str, gpu_block_size: int, gpu_blocks_per_file: int,
                 tp_size: int, pp_size: int, pcp_size: int, rank: int, dtype: str):
        # <root>/<model>/block_size_<n>_blocks_per_file_<m>/tp_<t>_pp_size_<p>_pcp_size_<c>/rank_<r>/<dtype>
        parts = [
            str(root_dir), str(model_name),
            "block_size_%d_blocks_per_file_%d" % (gpu_block_size, gpu_blocks_per_file),
            "tp_%d_pp_size_%d_pcp_size_%d" % (tp_size, pp_size, pcp_size),
            "rank_%d" % rank, str(dtype),
        ]
        self.base_path = "/".join(parts)

    def get_file_name(self, block_hash) -> str:
        """<base>/<hhh>/<hh>/<016x>.bin; bytes hashes are big-endian, only the low 64 bits are used."""
        value = int.from_bytes(block_hash, "big") if isinstance(block_hash, (bytes, bytearray)) else block_hash
        if not isinstance(value, int):
            raise TypeError("block_hash must be int or bytes")
        name = format(value & 0xFFFFFFFFFFFFFFFF, "016x")
        return "/".join((self.base_path, name[0:3], name[3:5], name + ".bin"))


def hashes_low64(block_hashes) -> "np.ndarray":
    """The low 64 bits of each block hash as a uint64 array — what get_file_name keeps of it (bytes hashes are
    big-endian).  Equal-length bytes hashes (what vLLM hands over) are converted without a Python loop."""
    import numpy as np
    hs = block_hashes if isinstance(block_hashes, (list, tuple)) else list(block_hashes)
    n = len(hs)
    if n == 0:
        return np.empty(0, dtype=np.uint64)
    first = hs[0]
    if isinstance(first, (bytes, bytearray)):
        width = len(first)
        raw = b"".join(hs)
        if width >= 8 and len(raw) == n * width:  # all the same length (else fall through to the loop)
            tail = np.frombuffer(raw, dtype=np.uint8).reshape(n, width)[:, width - 8:]
            return np.ascontiguousarray(tail).view(">u8").reshape(n).astype(np.uint64)
    out = np.empty(n, dtype=np.uint64)
    for i, h in enumerate(hs):
        v = int.from_bytes(h, "big") if isinstance(h, (bytes, bytearray)) else int(h)
        out[i] = v & 0xFFFFFFFFFFFFFFFF
    return out
