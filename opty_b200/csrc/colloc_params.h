// Kernel parameter block shared by the generated modules (colloc_kernel.cuh)
// and the host runtime (runtime.cu).
#pragma once

struct OptyParams {
  double* traj;   // [R + D][ldt] trajectory matrix + derived rows
  double* con;    // [M][ldc] residuals, eom-major
  double* jac;    // [nodes][K] partials, node-major
  long long ldt;
  long long ldc;
  int n_nodes;    // constraint nodes in this launch (N - 1 or a shard of them)
  int n_cols;     // valid trajectory columns (n_nodes + 1)
};
