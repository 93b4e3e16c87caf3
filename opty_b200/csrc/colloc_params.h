// Kernel parameter block shared by the generated modules (colloc_kernel.cuh)
// and the host runtime (runtime.cu).
#pragma once

struct OptyParams {
  double* traj;   // [R + D][ldt] trajectory matrix + derived rows
  double* tiled;  // direct-input modules: the same values tile by tile, [tiles][R + D][32*W + 2]: node tile,
                  // row, node inside the tile plus the two columns that follow it -- every row of a block's
                  // slice at a compile-time offset from one base pointer
  double* con;    // [M][ldc] residuals, eom-major
  double* jac;    // [nodes][K] partials, node-major
  long long ldt;
  long long ldc;
  int n_nodes;    // constraint nodes in this launch (N - 1 or a shard of them)
  int n_cols;     // valid trajectory columns (n_nodes + 1)
  int n_tiles;    // node tiles of 32*W nodes (persistent kernel)
  int* work;      // persistent kernel: next tile per group [groups of the module] + departure counter;
                  // row-stationary kernel: launch number that last claimed each slot [slots]
  int epoch;      // row-stationary kernel: number of this launch on the handle (1, 2, ...)
  const double* cvals;  // row-stationary kernel: values of the constant column runs (tail of the invariants table)
  unsigned long long* ready;  // row-stationary kernel with fused pre-pass: (node, case) pairs done per node tile
                              // [n_tiles] and in total [1], summed over all launches of the handle
};
