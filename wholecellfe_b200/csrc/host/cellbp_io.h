// cellbp_io.h -- mesh input and .vtk frame output in the reference's formats (cellBP/iomesh.cpp:491-565, :730-836, :844-1018).
#pragma once
#include <string>
#include <vector>

namespace cellbp {

// "edited COMSOL" mesh: Nodes / N / x y ... / Elements / M / i j k (0-based); z is forced to 0 (iomesh.cpp:523)
void read_ecomsol(const std::string &path, std::vector<double> &xyz, std::vector<int> &tri);

// cell frame (iomesh.cpp:844-1018): POINTS/CELLS/CELL_TYPES, SCALARS concentrations float 4 = (c_A, theta_e, c_C, F_int),
// VECTORS fa_u; default ostream precision.  kint < 0 selects the MD force law for F_int (kMD iomesh.cpp:924-956).
void write_vtk_cell(const std::string &path, double time, const std::vector<double> &xyz, const std::vector<int> &tri,
                    const std::vector<double> &cA, const std::vector<double> &cC, const std::vector<double> &fa_u,
                    const std::vector<double> &ea, double kint);
// substrate frame (iomesh.cpp:730-836): std::fixed after the POINTS header; SCALARS theta_e float 1; VECTORS fa_u
void write_vtk_subs(const std::string &path, double time, const std::vector<double> &xyz, const std::vector<int> &tri,
                    const std::vector<double> &ea, const std::vector<double> &fa_u_rows);

}   // namespace cellbp
