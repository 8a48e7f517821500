#include "cellbp_io.h"
#include <cmath>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace cellbp {

void read_ecomsol(const std::string &path, std::vector<double> &xyz, std::vector<int> &tri) {
    std::ifstream f(path);
    if (!f.is_open()) throw std::runtime_error("cannot open mesh file " + path);
    xyz.clear(); tri.clear();
    std::string line;
    bool have_nodes = false, have_elems = false;
    while (std::getline(f, line)) {
        if (line.find("Nodes") != std::string::npos) {
            std::getline(f, line);
            const int n = std::atoi(line.c_str());
            for (int i = 0; i < n; i++) {
                if (!std::getline(f, line)) throw std::runtime_error(path + ": truncated Nodes block");
                std::stringstream ss(line);
                double x = 0, y = 0; ss >> x >> y;
                xyz.push_back(x); xyz.push_back(y); xyz.push_back(0.0);              // planar problems: z := 0
            }
            have_nodes = true;
        } else if (line.find("Elements") != std::string::npos) {
            std::getline(f, line);
            const int m = std::atoi(line.c_str());
            for (int i = 0; i < m; i++) {
                if (!std::getline(f, line)) throw std::runtime_error(path + ": truncated Elements block");
                std::stringstream ss(line);
                int a = 0, b = 0, c = 0; ss >> a >> b >> c;
                tri.push_back(a); tri.push_back(b); tri.push_back(c);
            }
            have_elems = true;
        }
    }
    if (!have_nodes || !have_elems) throw std::runtime_error(path + ": not an ecomsol mesh (Nodes / Elements blocks missing)");
}

namespace {
// theta_e = sqrt(det((I - 2 ea)^-1)) per node (iomesh.cpp:789-802, :905-919), in the reference's own operation order: it goes
// through Eigen 3.3.4's fixed-size 3x3 inverse (cofactors * 1/det, det = c00 m00 + (c10 m10 + c20 m20):
// Eigen/src/LU/InverseImpl.h compute_inverse_size3_helper) and then Eigen's 3x3 determinant of that inverse
// (Eigen/src/LU/Determinant.h bruteforce_det3_helper) -- the printed 6 significant digits then agree in the last place too
double theta_e_of(const double *e) {
    double m[3][3] = {{1.0 - 2.0 * e[0], 0.0 - 2.0 * e[3], 0.0 - 2.0 * e[4]},
                      {0.0 - 2.0 * e[3], 1.0 - 2.0 * e[1], 0.0 - 2.0 * e[5]},
                      {0.0 - 2.0 * e[4], 0.0 - 2.0 * e[5], 1.0 - 2.0 * e[2]}};
    auto cof = [&](int i, int j) {
        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
        return m[i1][j1] * m[i2][j2] - m[i1][j2] * m[i2][j1];
    };
    const double c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
    const double det = c00 * m[0][0] + (c10 * m[1][0] + c20 * m[2][0]);
    const double invdet = 1.0 / det;
    double b[3][3];
    b[0][0] = c00 * invdet; b[0][1] = c10 * invdet; b[0][2] = c20 * invdet;
    b[1][0] = cof(0, 1) * invdet; b[1][1] = cof(1, 1) * invdet; b[1][2] = cof(2, 1) * invdet;
    b[2][0] = cof(0, 2) * invdet; b[2][1] = cof(1, 2) * invdet; b[2][2] = cof(2, 2) * invdet;
    auto h = [&](int a, int bb, int c) { return b[0][a] * (b[1][bb] * b[2][c] - b[1][c] * b[2][bb]); };
    return std::sqrt(h(0, 1, 2) - h(1, 0, 2) + h(2, 0, 1));
}
void header(std::ofstream &o, double time, size_t nn) {
    o << "# vtk DataFile Version 2.0\nCellBP Time: " << time << "\nASCII\nDATASET UNSTRUCTURED_GRID\n";
    o << "POINTS " << nn << " double\n";
}
void cells(std::ofstream &o, const std::vector<int> &tri) {
    const size_t ne = tri.size() / 3;
    o << "CELLS " << ne << " " << ne * 4 << "\n";
    for (size_t e = 0; e < ne; e++) o << "3 " << tri[3 * e] << " " << tri[3 * e + 1] << " " << tri[3 * e + 2] << "\n";
    o << "CELL_TYPES " << ne << "\n";
    for (size_t e = 0; e < ne; e++) o << "5\n";
}
const double MD_PX[6] = {0., 2.51250583e-03, 4.40265201e-03, 5.91408228e-03, 11.76817302e-03, 17.001e-03};
const double MD_PY[6] = {-7.15514497, 71.80501308, 356.62420075, 356.96656556, 126.21811585, 528.40481124};
const double MD_K[6] = {31426.8556, 99485.86235, 67080.67086, -7908.39043, 21978.66923, 76858.39733};
}   // namespace

void write_vtk_cell(const std::string &path, double time, const std::vector<double> &xyz, const std::vector<int> &tri,
                    const std::vector<double> &cA, const std::vector<double> &cC, const std::vector<double> &fa_u,
                    const std::vector<double> &ea, double kint) {
    std::ofstream o(path);
    if (!o) throw std::runtime_error("Unable to open output file.");                   // iomesh.cpp:692-695
    const size_t nn = xyz.size() / 3;
    header(o, time, nn);
    for (size_t i = 0; i < nn; i++) o << xyz[3 * i] << " " << xyz[3 * i + 1] << " " << xyz[3 * i + 2] << "\n";
    cells(o, tri);
    o << "POINT_DATA " << nn << "\n" << "SCALARS concentrations float 4\n" << "LOOKUP_TABLE default\n";
    for (size_t i = 0; i < nn; i++) {
        const double *u = &fa_u[3 * i];
        double fint;
        if (kint < 0) {                                                                // kMD iomesh.cpp:924-956
            const double D = std::sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
            fint = 0;
            for (int j = 0; j < 5; j++) {
                if (D < MD_PX[j + 1] && D >= MD_PX[j]) fint = MD_PY[j] + MD_K[j] * (D - MD_PX[j]);
                else if (D >= MD_PX[5]) fint = MD_PY[5] + MD_K[5] * (D - MD_PX[5]);
            }
        } else {
            const double fx = kint * u[0], fy = kint * u[1];                           // iomesh.cpp:923-925
            fint = std::sqrt(fx * fx + fy * fy);
        }
        o << cA[i] << " " << theta_e_of(&ea[6 * i]) << " " << cC[i] << " " << fint << "\n";
    }
    o << "VECTORS fa_u float\n";
    for (size_t i = 0; i < nn; i++) o << fa_u[3 * i] << " " << fa_u[3 * i + 1] << " " << fa_u[3 * i + 2] << "\n";
}

void write_vtk_subs(const std::string &path, double time, const std::vector<double> &xyz, const std::vector<int> &tri,
                    const std::vector<double> &ea, const std::vector<double> &fa_u_rows) {
    std::ofstream o(path);
    if (!o) throw std::runtime_error("Unable to open output file.");
    const size_t nn = xyz.size() / 3;
    header(o, time, nn);
    o << std::fixed;                                                                   // iomesh.cpp:739: after the POINTS header
    for (size_t i = 0; i < nn; i++) o << xyz[3 * i] << " " << xyz[3 * i + 1] << " " << xyz[3 * i + 2] << "\n";
    cells(o, tri);
    o << "POINT_DATA " << nn << "\n" << "SCALARS theta_e float 1\n" << "LOOKUP_TABLE default\n";
    for (size_t i = 0; i < nn; i++) o << theta_e_of(&ea[6 * i]) << "\n";
    o << "VECTORS fa_u float\n";
    // the reference hands the substrate writer a fa_u vector that is not the substrate's (quirk q6: the cell's at frame 0,
    // an out-of-bounds read later); here: the rows given, zero-padded
    for (size_t i = 0; i < nn; i++) {
        if (3 * i + 2 < fa_u_rows.size()) o << fa_u_rows[3 * i] << " " << fa_u_rows[3 * i + 1] << " " << fa_u_rows[3 * i + 2] << "\n";
        else o << 0.0 << " " << 0.0 << " " << 0.0 << "\n";
    }
}

}   // namespace cellbp
