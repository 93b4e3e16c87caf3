"""Kernel-option variants exercised by the GPU suite.  Kept in one place so
that ``__graft_entry__.build()`` can compile exactly these modules into the
in-tree cache before the snapshot travels to the GPU box."""

# tests/test_gpu_parity.py::test_config2_kernel_variants_agree_bitwise
# (every variant additionally gets fmad=False, reassociate=False)
CONFIG2_VARIANTS = [
    {'groups': 1},
    {'schedule': False, 'groups': 8},
    {'groups': 11, 'tile_cols': 14, 'warps_per_block': 4,
     'min_blocks_per_sm': 2, 'live_budget': 24},
    {'tma_store': False, 'tma_load': False, 'groups': 5},
    {'d2h_skip_constants': False, 'groups': 3, 'out_ring': 3,
     'tile_bufs': 2, 'min_blocks_per_sm': 3, 'live_budget': 100},
    {'pre_pass': False, 'groups': 8, 'warps_per_block': 1,
     'min_blocks_per_sm': 8},
    # direct input loads, one staging buffer per warp holding two rows
    {'groups': 8, 'tma_load': 'direct', 'tile_cols': 92, 'tile_bufs': 1,
     'warps_per_block': 1, 'min_blocks_per_sm': 8},
    # the problem compiled as three modules (parallel nvcc runs)
    {'compile_shards': 3, 'groups': 8, 'out_ring': 2},
    # 8-warp blocks with direct input loads
    {'groups': 8, 'warps_per_block': 8, 'min_blocks_per_sm': 1,
     'tma_load': 'direct', 'compile_shards': 2},
    # code-stationary persistent kernel (TMA input / direct input, two
    # modules)
    {'persistent': True, 'groups': 11},
    {'persistent': True, 'groups': 6, 'tma_load': 'direct',
     'compile_shards': 2, 'tile_bufs': 2},
    # the grid kernel with automatic grouping (the default is the
    # row-stationary kernel at this problem)
    {'persistent': False},
    # row-stationary kernel: narrow tiles in two staging buffers (an item has
    # an odd number of phases: the buffers swap), the pre-pass fused into the
    # main kernel, constant rows as ordinary groups, plain input loads
    {'persistent': 'stationary', 'tile_bufs': 2, 'tile_cols': 22},
    {'persistent': 'stationary', 'tile_bufs': 1, 'fused_pre': True},
    {'persistent': 'stationary', 'tile_bufs': 1, 'const_rows': False,
     'warps_per_block': 4, 'store_hint': 0},
]
BITWISE = {'fmad': False, 'reassociate': False}

# tests/test_gpu_parity.py::test_config4_standin_against_oracle
CONFIG4_VARIANTS = [
    {},
    # odd P (27): a staging buffer holds two equation rows; here cut further
    {'tile_cols': 20, 'groups': 4, 'tile_bufs': 2},
    {'schedule': False, 'tma_load': 'direct'},
    # persistent kernel on a backward-Euler problem with a known trajectory,
    # free parameters and a free time interval (invariants change per call)
    {'persistent': True, 'groups': 4, 'warps_per_block': 4,
     'min_blocks_per_sm': 2},
    # row-stationary kernel forced onto an odd P: two equations per group,
    # no constant rows, invariants that change with every call
    {'persistent': 'stationary', 'tile_bufs': 1},
]
CONFIG4_IDS = ['default', 'narrow_tiles', 'unscheduled', 'persistent',
               'stationary']

# tests/test_gpu_parity.py::test_node_range_shards_reproduce_the_whole
CONFIG2_SHARD_BOUNDS = [0, 1, 2500, 7001, 9999]
