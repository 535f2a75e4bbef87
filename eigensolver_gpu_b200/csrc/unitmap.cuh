// eigb200 -- work decomposition of the symv/hemv tile engine (csrc/sytrd.cu): strip length heuristic, the unit map
// (unit id -> tiles) and the per-column descriptor.  Pure integer/scalar logic, compiled for the device by nvcc and
// for the host by the CPU unit test (tests/csrc/unitmap_host.cpp, tests/test_unitmap_host.py).
#pragma once
#include <math.h>
#ifdef __CUDACC__
#define EIGB_HD __host__ __device__ __forceinline__
#else
#define EIGB_HD inline
#endif

namespace eigb200 {
namespace tile {

constexpr int TB = 64;        // symv/hemv tile edge
constexpr int MAXBANDS = 256; // max number of strip bands (tile rows / strip length), enforced by strip_len()

// per-column descriptor, see compute_desc()
struct ColDesc {
  int j, Tn, C, rcpC, KB, NF, total;   // order, tile rows, strip length (+ 2^16 reciprocal), unit map (P == 1)
  int R;                               // rows per CTA in the phase A that follows
  int ndj, nsj;                        // partial-sum slots of the last row (direct / total)
};


// tiles per strip chunk for an order-n product on G CTAs: aim at >= upc units per CTA, at most 8 tiles per unit
EIGB_HD int strip_len(int n, int G, int P, int upc) {
  const int Tn = (n + TB - 1) / TB;
  const int cmax = upc >> 8 ? upc >> 8 : 8;     // (tuning) bits 8.. of upc override the maximum strip length
  upc &= 255;
  int c = (Tn * (Tn - 1) / 2) / (upc * G * P);
  if (c < 1) c = 1;
  if (c > cmax) c = cmax;
  while ((Tn - 1) / c > MAXBANDS) ++c;
  return c;
}

// Everything a CTA needs to know about the product of order j (panel column c).  It is used by phase B(c) and by
// the phase A that follows (c-1) and is derived one column ahead by a single thread that would otherwise idle
// (the producer warp, after its last tile): integer divisions and square roots cost ~25 dependent instructions
// each, and 17 warps repeating them on the critical path of every column was a measurable part of it.
EIGB_HD void compute_desc(ColDesc& d, int j, int G, int P, int upc, const unsigned char* ctab = nullptr) {
  d.j = j;
  if (j <= 0) { d.Tn = 0; d.C = 1; d.rcpC = 65536; d.KB = 0; d.NF = 0; d.total = 0; d.R = 0; d.ndj = 0; d.nsj = 0; return; }
  const int Tn = (j + TB - 1) / TB;
  // strip length: from the host-built table (build_strip_table) when there is one, else the closed-form heuristic
  const int C = ctab != nullptr ? (int)ctab[Tn] : strip_len(j, G, P, upc);
  d.Tn = Tn; d.C = C; d.rcpC = (65536 + C - 1) / C;
  d.KB = (Tn - 1) / C;
  d.NF = d.KB * Tn - C * (d.KB * (d.KB + 1) / 2);
  d.total = d.NF + Tn;                 // P == 1 (P > 1: engine_prepare tabulates the owned units)
  int R = (j + G - 1) / G;
  d.R = (R + 7) & ~7;
  const int Ij = (j - 1) / TB;         // tile row of the last row (row j' = j-1 of the following phase A)
  d.ndj = Tn - (Ij + 1);
  d.nsj = d.ndj + Ij / C + 1;
}


// Units of the product of order n with strip length C on P ranks (rank owns tile columns J = rank (mod P)):
//   F(k, J)  band k (tile rows [kC, kC+C)), tile column J >= (k+1)C : C off-diagonal tiles
//   D(J)     tile column J: off-diagonal tiles of the partial band floor(J/C) + the diagonal tile
// Unit ids: F units band by band, then (P == 1) the D units by decreasing size.
struct UnitMap {
  int Tn, C, rcpC, rank, P, TnO, KB, NF, total;
  int keepI;            // tiles of tile rows < keepI are loaded with L2 evict_last, the others with evict_first (0: no hints)
  const int* bstart;    // P > 1: bstart[k] = number of F units in bands < k (shared memory, KB+1 entries)
  // owned tile columns are J = rank + P*jj, jj = 0..TnO-1
  EIGB_HD int first_owned_at_least(int Jmin) const {   // smallest jj with rank + P*jj >= Jmin
    const int d = Jmin - rank;
    return d <= 0 ? 0 : (d + P - 1) / P;
  }
  EIGB_HD int band_count(int k) const {                // F units of band k
    const int jj0 = first_owned_at_least((k + 1) * C);
    return TnO - jj0 > 0 ? TnO - jj0 : 0;
  }
  // P == 1: F units before band k = k*Tn - C*k*(k+1)/2
  EIGB_HD int prefix1(int k) const { return k * Tn - C * (k * (k + 1) / 2); }
  // number of bands that have F units at all: (k+1)*C < Tn
  EIGB_HD static int num_bands(int Tn_, int C_) { return Tn_ > 0 ? (Tn_ - 1) / C_ : 0; }
  EIGB_HD void init(int n, int C_, int rank_, int P_, const int* bstart_) {
    Tn = (n + TB - 1) / TB; C = C_; rcpC = (65536 + C_ - 1) / C_; rank = rank_; P = P_; bstart = bstart_; keepI = 0;
    TnO = rank < Tn ? (Tn - rank + P - 1) / P : 0;
    KB = num_bands(Tn, C);
    NF = (P == 1) ? prefix1(KB) : (KB > 0 ? bstart[KB] : 0);
    total = NF + TnO;
  }
  // tiles of a unit: off-diagonal (I, J) for I in [I0, I1), then the diagonal tile (J, J) when has_diag
  EIGB_HD void decode(int unit, int& J, int& I0, int& I1, bool& has_diag) const {
    if (unit < NF) {
      int k;
      if (P == 1) {
        const double bq = (double)Tn - 0.5 * (double)C;
        k = (int)((bq - sqrt(fmax(bq * bq - 2.0 * (double)C * (double)unit, 0.0))) / (double)C);
        if (k < 0) k = 0;
        if (k > KB - 1) k = KB - 1;
        while (k > 0 && prefix1(k) > unit) --k;
        while (k + 1 < KB && prefix1(k + 1) <= unit) ++k;
        J = (k + 1) * C + (unit - prefix1(k));
      } else {
        int lo = 0, hi = KB - 1;                 // largest k with bstart[k] <= unit
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (bstart[mid] <= unit) lo = mid; else hi = mid - 1; }
        k = lo;
        J = rank + P * (first_owned_at_least((k + 1) * C) + (unit - bstart[k]));
      }
      I0 = k * C; I1 = k * C + C; has_diag = false;
    } else {
      int q = unit - NF;
      if (P == 1) {
        // decreasing size: a D unit has (J mod C) + 1 tiles; residue classes from the largest down
        int s = (C < Tn ? C : Tn) - 1;
        for (; s > 0; --s) {
          const int cnt = (Tn - 1 - s) / C + 1;
          if (q < cnt) break;
          q -= cnt;
        }
        J = s + C * q;
      } else {
        J = rank + P * q;
      }
      I0 = (J / C) * C; I1 = J; has_diag = true;
    }
  }
};


// Host side: strip length per number of tile rows, chosen by replaying the unit queue.  The F units of a product all have
// C tiles, so after them the CTAs stand at q or q+1 units; the D units (1..C tiles, in queue order) then go to the
// earliest free CTA.  A unit costs its tiles plus `ov` tile times (descriptor, strip end).  The closed-form heuristic
// (strip_len) leaves up to a unit of imbalance at the end of a column -- 10 us of a 50 us column on two ranks.
inline void build_strip_table(int G, int P, int Tnmax, int cmax, double ov, unsigned char* out /* Tnmax + 1 entries */) {
  out[0] = 1;
  double* fin = new double[G];
  for (int Tn = 1; Tn <= Tnmax; ++Tn) {
    int bestC = 1; double best = 1e300;
    for (int C = 1; C <= cmax; ++C) {
      if ((Tn - 1) / C > MAXBANDS) continue;
      double worst = 0.0;
      for (int rank = 0; rank < P; ++rank) {
        UnitMap um; um.Tn = Tn; um.C = C; um.rank = rank; um.P = P;
        um.TnO = rank < Tn ? (Tn - rank + P - 1) / P : 0;
        const int KB = UnitMap::num_bands(Tn, C);
        long long NF = 0;
        for (int k = 0; k < KB; ++k) NF += um.band_count(k);
        const double u = C + ov;
        const long long q = NF / G; const int rem = (int)(NF % G);
        for (int i = 0; i < G; ++i) fin[i] = (double)(q + (i < rem ? 1 : 0)) * u;
        // D units in queue order: P == 1 by decreasing size, P > 1 by increasing tile column
        auto place = [&](int tiles) {
          int im = 0;
          for (int i = 1; i < G; ++i) if (fin[i] < fin[im]) im = i;
          fin[im] += tiles + ov;
        };
        if (P == 1) {
          for (int s = (C < Tn ? C : Tn) - 1; s >= 0; --s)
            for (int J = s; J < Tn; J += C) place(s + 1);
        } else {
          for (int J = rank; J < Tn; J += P) place(J % C + 1);
        }
        double mk = 0.0;
        for (int i = 0; i < G; ++i) if (fin[i] > mk) mk = fin[i];
        if (mk > worst) worst = mk;
      }
      if (worst <= best) { best = worst; bestC = C; }      // ties: the longer strip
    }
    out[Tn] = (unsigned char)bestC;
  }
  delete[] fin;
}

}  // namespace tile
}  // namespace eigb200
