// glc_tables_host.h -- host-side pre-processing of the tabulated inputs, as the reference's readers do once at
// start-up (cieFileReadFile, cooling/cooling_function/CIE_file.F90:627-659; virial_density_contrast.F90:386-414;
// exponential_disk.F90:675-733).  Pure C++ (no CUDA calls): shared by glc_api.cu and by tests/emu.
#pragma once

#include <vector>

#include "glc_common.cuh"

namespace glc {

struct PreparedTable {
    std::vector<double> x0, x1, v;
    int is_log = 0, first_zero = 0;
    double first_nonzero = 0.0, zmin = 0.0, zmax = 0.0, tmin = 0.0, tmax = 0.0;
    double ln0 = 0.0, inv_dln = 0.0;  // uniform-in-log lattices
};

inline int prepare_table(int id, int n0, int n1, const double *x0, const double *x1, const double *values,
                         PreparedTable &t) {
    if (id < 0 || id >= GLC_NTABLES || n0 < 2 || n1 < 1 || !x0 || !values) return -1;
    t.x0.assign(x0, x0 + n0);
    t.v.assign(values, values + (size_t)n0 * n1);
    if (x1) t.x1.assign(x1, x1 + n1);
    t.zmin = t.x0.front();
    t.zmax = t.x0.back();
    t.tmin = x1 ? t.x1.front() : 0.0;
    t.tmax = x1 ? t.x1.back() : 0.0;
    if (id == GLC_TABLE_COOLING_FUNCTION || id == GLC_TABLE_ELECTRON_FRACTION) {
        if (!x1 || n1 < 2) return -1;
        t.is_log = 1;
        for (double v : t.v)
            if (!(v > 0.0)) t.is_log = 0;
        if (t.is_log) {
            t.first_zero = (t.x0[0] == 0.0);
            if (t.first_zero) t.first_nonzero = t.x0[1];
            for (auto &z : t.x0) z = (z > 0.0) ? dm_log(z) : -999.0;
            for (auto &T : t.x1) T = dm_log(T);
            for (auto &v : t.v) v = dm_log(v);
        }
    } else if (id == GLC_TABLE_HALO_MEAN_DENSITY) {
        if (n1 != 2) return -1;
        // ln t on the device; the grid must be log-uniform
        for (auto &x : t.x0) x = dm_log(x);
        t.ln0 = t.x0[0];
        t.inv_dln = (double)(n0 - 1) / (t.x0[n0 - 1] - t.x0[0]);
    } else if (id == GLC_TABLE_DISK_ROTATION_CURVE) {
        if (n1 != 1) return -1;
        t.ln0 = dm_log(t.x0[0]);
        t.inv_dln = (double)(n0 - 1) / (dm_log(t.x0[n0 - 1]) - dm_log(t.x0[0]));
    } else if (id == GLC_TABLE_ADAF) {
        if (n1 != 2) return -1;
        // table1DLogarithmicLinear keeps its abscissae as ln x (objects/tables/_module.F90:1451-1468)
        for (auto &x : t.x0) x = dm_log(x);
        t.ln0 = t.x0[0];
        t.inv_dln = (double)(n0 - 1) / (t.x0[n0 - 1] - t.x0[0]);
    }
    return 0;
}

// lattice values of a fastExponentiator table: x_k = rangeMin + dx k (k < n-1), x_{n-1} = rangeMax, exactly the
// abscissae table1DLinearLinear would be populated on (objects/tables/_module.F90:1237-1405)
inline std::vector<double> build_pow_table(double rangeMin, double rangeMax, double exponent, double density) {
    const int n = (int)((rangeMax - rangeMin) * density) + 1;
    const double dx = (rangeMax - rangeMin) / (double)(n - 1);
    std::vector<double> t((size_t)n);
    for (int k = 0; k < n; k++) t[k] = dm_pow((k == n - 1) ? rangeMax : rangeMin + dx * (double)k, exponent);
    return t;
}

// Inverse tabulation used by nfwRadiusFromSpecificAngularMomentum (mass_distributions/spherical/NFW.F90:589-639): abscissae on
// the absolute octave lattice x_k = 2^(k/30) (countRadiiPerOctave = 30, NFW.F90:94; numerical/ranges.F90 Lattice_Value),
// values sqrt(massEnclosedScaleFree(x) x) with the 4 pi of NFW.F90:568-570.  The reference grows the lattice octave by octave
// until it brackets the request; 80 octaves about unity cover every physical case and, the lattice being absolute, hold the
// very points the reference would compute.
constexpr int kNfwJPointsPerOctave = 30, kNfwJOctaves = 40;
inline void build_nfw_j_table(std::vector<double> &xs, std::vector<double> &js) {
    const int n = 2 * kNfwJOctaves * kNfwJPointsPerOctave + 1;
    const double ln2 = 0.69314718055994530942, pi = 3.14159265358979323846;
    xs.resize(n);
    js.resize(n);
    for (int i = 0; i < n; i++) {
        const int k = i - kNfwJOctaves * kNfwJPointsPerOctave;
        const double x = (k % kNfwJPointsPerOctave == 0) ? dm_scale2(1.0, k / kNfwJPointsPerOctave)
                                                         : dm_exp(((double)k / (double)kNfwJPointsPerOctave) * ln2);
        double m;
        if (x == 1.0)
            m = dm_log(2.0) - 0.5;
        else if (x >= 1.0e-6)
            m = dm_log(1.0 + x) - x / (1.0 + x);
        else
            m = x * x * (0.5 + x * (-2.0 / 3.0 + x * (0.75 + x * (-0.8))));
        m = 4.0 * pi * m;
        xs[i] = x;
        js[i] = sqrt(m * x);
    }
}

// bin edges of mergerTreeEvolveProfilerSimple (simple.F90:128-147): Make_Range(min, max, n, logarithmic) with
// n = int(log10(max / min) * pointsPerDecade) + 1, i.e. exp of a linear range of the logarithms
inline int build_profile_edges(const glc_params &P, double *edges) {
    int n = (int)(log10(P.profilerTimeStepMaximum / P.profilerTimeStepMinimum) * (double)P.profilerTimeStepPointsPerDecade) + 1;
    if (n < 2) n = 2;
    if (n > GLC_PROFILE_BINS) n = GLC_PROFILE_BINS;
    const double l0 = dm_log(P.profilerTimeStepMinimum), l1 = dm_log(P.profilerTimeStepMaximum);
    for (int i = 0; i < n; i++) edges[i] = dm_exp(l0 + (l1 - l0) * (double)i / (double)(n - 1));
    for (int i = n; i < GLC_PROFILE_BINS; i++) edges[i] = 0.0;
    return n;
}

inline void pow_table_spacing(double rangeMin, double rangeMax, int n, double &dx, double &inverseDx) {
    dx = (rangeMax - rangeMin) / (double)(n - 1);
    inverseDx = 1.0 / ((rangeMin + dx) - rangeMin);
}

inline void install_table(DeviceTables &T, int id, const PreparedTable &t, const DeviceTable2D &d) {
    if (id == GLC_TABLE_COOLING_FUNCTION) {
        T.cooling = d;
        T.cooling_log = t.is_log;
        T.cooling_first_z_zero = t.first_zero;
        T.cooling_first_nonzero_z = t.first_nonzero;
        T.cooling_z_min = t.zmin;
        T.cooling_z_max = t.zmax;
        T.cooling_t_min = t.tmin;
        T.cooling_t_max = t.tmax;
    } else if (id == GLC_TABLE_ELECTRON_FRACTION) {
        T.electron = d;
        T.electron_log = t.is_log;
        T.electron_first_z_zero = t.first_zero;
        T.electron_first_nonzero_z = t.first_nonzero;
        T.electron_z_min = t.zmin;
        T.electron_z_max = t.zmax;
        T.electron_t_min = t.tmin;
        T.electron_t_max = t.tmax;
    } else if (id == GLC_TABLE_HALO_MEAN_DENSITY) {
        T.density = d;
        T.density_lnt0 = t.ln0;
        T.density_inv_dlnt = t.inv_dln;
    } else if (id == GLC_TABLE_DISK_ROTATION_CURVE) {
        T.diskrc = d;
        T.diskrc_lnx0 = t.ln0;
        T.diskrc_inv_dlnx = t.inv_dln;
    } else if (id == GLC_TABLE_ADAF) {
        T.adaf = d;
        T.adaf_inv_dlnx = t.inv_dln;
    }
}

}  // namespace glc
