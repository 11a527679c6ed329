// lattice.h -- integer FCC geometry shared by host and device code.
//
// Replaces the floating-point geometry of the reference's cfg::Config
// (/root/reference/lmc/cfg/src/Config.cpp: GenerateFCC :1060-1095, ReassignLatticeVector :466-552,
// UpdateNeighbors :955-1045) with exact integer arithmetic on half-lattice-constant coordinates:
// an FCC site is a point (X,Y,Z), 0 <= X < 2fx etc., with X+Y+Z even.
//
// Two lattice-id orders of the reference are supported natively:
//   LMC_ORDER_GENERATE   id = ((k*fy + j)*fx + i)*4 + b      (cell (i,j,k), basis b; Config.cpp:1073-1090)
//   LMC_ORDER_REASSIGNED id = X*(2*fy*fz) + Y*fz + Z/2        (sorted by x, then y, then z; Config.cpp:466-478)
//
// Device storage ("padded layout"): occupancy lives in HBM as one compact uint8 code per site in a
// 3-D array indexed by half-unit coordinates with the unused parity squeezed out of z and a periodic
// halo of 3 half-units on every face, so that every neighbourhood gather is `base + constant offset`
// with no wrap arithmetic:   index = ((X+3)*NY + (Y+3))*NZ + ((Z+4)>>1),  NY = 2fy+6, NZ = fz+4.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define LMC_HD __host__ __device__ __forceinline__
#else
#define LMC_HD inline
#endif

namespace lmc {

enum IdOrder : int32_t { LMC_ORDER_GENERATE = 0, LMC_ORDER_REASSIGNED = 1 };

constexpr int kHalo = 3;        // half-units of periodic halo in x and y (max |offset| of any gather is 3)
constexpr int kHaloZ = 4;       // even halo in z so that the (Z+4)>>1 squeeze keeps parity

struct LatticeDesc {
  int32_t fx, fy, fz;           // supercell factors (conventional cells per axis)
  int32_t order;                // IdOrder
  int32_t nx, ny, nz;           // padded dims: nx = 2fx+6, ny = 2fy+6, nz = fz+4
  int64_t num_sites;            // 4*fx*fy*fz
  int64_t padded_size;          // nx*ny*nz

  LMC_HD void coords_of_id(int64_t id, int &X, int &Y, int &Z) const {
    if (order == LMC_ORDER_GENERATE) {
      const int b = static_cast<int>(id & 3);
      int64_t c = id >> 2;
      const int i = static_cast<int>(c % fx); c /= fx;
      const int j = static_cast<int>(c % fy);
      const int k = static_cast<int>(c / fy);
      X = 2 * i + ((b == 1) | (b == 2));
      Y = 2 * j + ((b == 1) | (b == 3));
      Z = 2 * k + ((b == 2) | (b == 3));
    } else {
      const int64_t per_x = 2LL * fy * fz;
      X = static_cast<int>(id / per_x);
      const int64_t r = id - X * per_x;
      Y = static_cast<int>(r / fz);
      const int zi = static_cast<int>(r - static_cast<int64_t>(Y) * fz);
      Z = 2 * zi + ((X + Y) & 1);
    }
  }
  // coordinates must already be wrapped into [0, 2f)
  LMC_HD int64_t id_of_coords(int X, int Y, int Z) const {
    if (order == LMC_ORDER_GENERATE) {
      const int xo = X & 1, yo = Y & 1, zo = Z & 1;
      const int b = (xo & yo) ? 1 : ((xo & zo) ? 2 : ((yo & zo) ? 3 : 0));
      return ((static_cast<int64_t>(Z >> 1) * fy + (Y >> 1)) * fx + (X >> 1)) * 4 + b;
    }
    return static_cast<int64_t>(X) * (2LL * fy * fz) + static_cast<int64_t>(Y) * fz + (Z >> 1);
  }
  LMC_HD int64_t padded_index(int X, int Y, int Z) const {   // X,Y in [-3, 2f+2], Z in [-4, 2f+3]
    return (static_cast<int64_t>(X + kHalo) * ny + (Y + kHalo)) * nz + ((Z + kHaloZ) >> 1);
  }
  LMC_HD int64_t padded_index_of_id(int64_t id) const {
    int X, Y, Z;
    coords_of_id(id, X, Y, Z);
    return padded_index(X, Y, Z);
  }
  // linear offset in the padded layout of a displacement (dx,dy,dz) from a site whose Z has parity zpar
  LMC_HD int32_t padded_delta(int dx, int dy, int dz, int zpar) const {
    // ((Z+dz+4)>>1) - ((Z+4)>>1) for Z of parity zpar; written so the shift never sees a negative operand
    const int dzi = ((zpar + dz + 8) >> 1) - ((zpar + 8) >> 1);
    return (dx * ny + dy) * nz + dzi;
  }
};

inline LatticeDesc make_lattice(int fx, int fy, int fz, int order) {
  LatticeDesc d{};
  d.fx = fx; d.fy = fy; d.fz = fz; d.order = order;
  d.nx = 2 * fx + 2 * kHalo; d.ny = 2 * fy + 2 * kHalo; d.nz = fz + kHaloZ;
  d.num_sites = 4LL * fx * fy * fz;
  d.padded_size = static_cast<int64_t>(d.nx) * d.ny * d.nz;
  return d;
}

LMC_HD int wrap_coord(int v, int period) {   // |v| < 2*period
  return v < 0 ? v + period : (v >= period ? v - period : v);
}

}  // namespace lmc
