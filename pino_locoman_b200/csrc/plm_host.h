// Host-side table builder: robot description + OCP options -> kinematic-tree tables, row/column layout of the
// OCP (x, p, g, J_g) and the per-node-type lookup tables the kernels scatter Jacobian entries with.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/pino_locoman_b200.h"
#include "plm_types.h"
#include "plm_qp_types.h"

namespace plm {

struct HostTables {
  PlmModel model;
  PlmLayout layout;
  std::vector<int16_t> lut;             // pool of per-type luts
  std::vector<PlmConstEntry> consts;    // pool of per-type constant entries
  std::vector<int32_t> pat_rows, pat_cols;   // COO pattern of J_g in value order
  std::vector<std::vector<std::vector<int>>> type_rowcols;   // [type][row] -> local columns (ascending)
  QpLayout qp;                          // QP solver layout
  std::vector<int16_t> qp_idx;          // pool of the per-type CSR/CSC index tables (+ flat column / row indices)
  std::vector<int32_t> qp_idx32;        // pool of the flat pointer tables
  std::string error;
};

// returns false and fills error on failure
bool build_tables(const plm_robot_desc& robot, const plm_ocp_desc& ocp, HostTables& out);

}  // namespace plm
