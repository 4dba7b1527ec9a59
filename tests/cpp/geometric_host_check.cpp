// Host-only check of the Geometric-level helpers of the C++ mirror (mimosa_b200/host/mimosa_b200.hpp): reads a pose
// sequence, runs KeyframeGate exactly as Geometric::updateMap would, and the degeneracy flags of getFactors on given
// localizabilities; prints the decisions for tests/test_geometric_host.py to compare with the oracle.  No GPU call.
//   stdin : trans_thresh rot_thresh_deg forced  R_B_L(9)  n  then n x (R(9) t(3))
//           then m  then m x (loc_rot_comp(3) loc_trans_comp(3) thresh_rot thresh_trans)
#include <cstdio>
#include <iostream>

#include "../../mimosa_b200/host/mimosa_b200.hpp"

int main() {
  using namespace mimosa_b200;
  float tt, rt;
  size_t forced, n;
  std::array<double, 9> RBL;
  std::cin >> tt >> rt >> forced;
  for (auto& v : RBL) std::cin >> v;
  std::cin >> n;
  KeyframeGate gate(tt, rt, forced, RBL);
  for (size_t i = 0; i < n; ++i) {
    Pose p;
    for (auto& v : p.R) std::cin >> v;
    for (auto& v : p.t) std::cin >> v;
    const bool up = gate.shouldUpdate(p);
    if (up) gate.addKeyframe(p);
    std::printf("%d\n", up ? 1 : 0);
  }
  size_t m;
  std::cin >> m;
  for (size_t i = 0; i < m; ++i) {
    mb_linearization lin{};
    RegistrationConfig cfg;
    for (int a = 0; a < 3; ++a) std::cin >> lin.loc_rot_comp[a];
    for (int a = 0; a < 3; ++a) std::cin >> lin.loc_trans_comp[a];
    std::cin >> cfg.degen_thresh_rot >> cfg.degen_thresh_trans;
    for (int a = 0; a < 9; ++a) lin.eigvec_rot[a] = 1 + a, lin.eigvec_trans[a] = 11 + a;
    const DegeneracyInfo d = degeneracyInfo(lin, cfg);
    for (int a = 0; a < 6; ++a) std::printf("%d ", (int)d.degen_directions[a]);
    std::printf("%g %g %g\n", d.eigenvectors_block_matrix[0], d.eigenvectors_block_matrix[6 * 3 + 3], d.eigenvectors_block_matrix[3]);
  }
  return 0;
}
