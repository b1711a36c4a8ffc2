"""trialign_b200 -- B200-native three-sequence alignment scoring (the TRIALIGN / PE / MAX7 hot path).

The directory name carries the reference repository's name and therefore hyphens; import it with
``importlib.import_module("hw-accelerator-three-sequence-alignment_b200")`` or through the repo-root shim
``trialign_b200.py``.  Everything here is a thin ctypes layer over ``lib/libtrialign.so`` (C ABI in
``include/trialign.h``); there is no Python or CPU compute path.
"""
from .trialign import (  # noqa: F401
    BACKEND_AUTO,
    BACKEND_GPU,
    BACKEND_GPU_FAST,
    BACKEND_GPU_LITERAL,
    PATH_NAMES,
    TRIPLE_RTL,
    TRIPLE_SUM_OF_PAIRS,
    Scoring,
    TriAlign,
    TriAlignError,
    build_library,
    device_count,
    format_alignment,
    apply_option_args,
    get_option,
    last_devices,
    last_flagged,
    last_segments,
    multi_plan,
    set_option,
    last_path,
    last_timing,
    launch_count,
    library,
    library_path,
    pack_batch_ram128,
    pack_seq_ram128,
    read_scores,
    read_seq_dat,
    synth,
    triple_indices,
    unpack_seq_ram128,
    synth_batch,
    synth_batch_cuda,
    version,
    write_scores,
    write_seq_dat,
    write_tb_init,
)
from .parallel import ColumnPipeline, SlabPipeline, col_range, make_pipeline, shard_range, slab_range  # noqa: E402,F401
