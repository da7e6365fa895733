// CSR operator type shared between csr.cu (SpMV) and laplacian.cu (device assembly).
#pragma once
#include "operator.hpp"

namespace pmgx
{
struct CsrOperator : pmgx_operator
{
  long long nnz = 0;
  bool has_ghost_cols = false;
  DevBuf<int32_t> row_ptr;  // n_owned + 1
  DevBuf<int32_t> off_diag; // n_owned: first ghost-column entry of each row (src/csr.hpp:118-121)
  DevBuf<int32_t> cols;
  DevBuf<double> values;
  DevBuf<int32_t> ghost_rows; // rows with at least one ghost-column entry
  int n_ghost_rows = 0;
  void apply(double* x, double* y) override;
  void finish_setup(); // extracts diag^-1 (src/csr.hpp:101-112)
};
} // namespace pmgx
