"""hypersonic-rANS on B200 — host-side Python mirror of the C-ABI in include/hsrans_b200.h.

The directory name carries a hyphen (it is the package the task names), so import it through
``__graft_entry__.load_package()`` or ``importlib`` under the module name ``hypersonic_rans_b200``.

Only plumbing lives here: a ctypes binding of ``libhsrans_b200.so`` (``capi``) and the codec registry that
mirrors the reference's ``_Codecs[]`` table (``codecs``, /root/reference/src/main.cpp:172-236). All decoding
happens in the hand-written sm_100a kernels under ``csrc/``; there is no Python or CPU decode path.
"""
from . import capi  # noqa: F401
from .capi import (  # noqa: F401
    HSR_RAW, HSR_BLOCK, HSR_MT, HSR_RAW32BLK, FAMILY_NAMES, HsrError, Block, PreparedStream, lib, lib_path, version, device_count,
    last_error, set_option, get_option, capacity, host_alloc, HostBuffer, decode, decode_batch, decode_mt_multi, mt_index,
    mt_partition, make_hist, synth_zipf, encode_mt, encode_mt_device, encode_mt_device_indexed, encode_mt_index_bound, encode_mt_bound, encode_mt_policy, encode_mt_policy_device, observe_hist_device, normalize_hist_device, make_hist_segments_device,
)
from .codecs import CODECS, Codec, find_codec  # noqa: F401
from .sharding import ShardPlan, plan_shards, assemble_on  # noqa: F401
