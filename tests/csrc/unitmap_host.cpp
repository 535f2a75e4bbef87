// Host build of the tile engine's work decomposition (csrc/unitmap.cuh) for CPU unit tests.
#include "unitmap.cuh"
#include <vector>
using namespace eigb200::tile;

// Enumerates all tiles of all units of the order-n product for rank `rank` of P: out[4*t] = {unit, I, J, is_diag};
// returns the number of tiles (or -1 if `cap` is too small).  G only enters through the strip length.
extern "C" int unitmap_enumerate(int n, int G, int P, int rank, int upc, int* out, int cap, int* info /* C, total, NF */) {
  const int C = strip_len(n, G, P, upc);
  const int Tn = (n + TB - 1) / TB;
  std::vector<int> bstart(MAXBANDS + 2, 0);
  UnitMap um;
  if (P > 1) {
    // the table the device builds in engine_prepare()
    UnitMap tmp; tmp.Tn = Tn; tmp.C = C; tmp.rank = rank; tmp.P = P;
    tmp.TnO = rank < Tn ? (Tn - rank + P - 1) / P : 0;
    const int KB = UnitMap::num_bands(Tn, C);
    int acc = 0;
    for (int k = 0; k < KB; ++k) { bstart[k] = acc; acc += tmp.band_count(k); }
    bstart[KB] = acc;
  }
  um.init(n, C, rank, P, bstart.data());
  info[0] = C; info[1] = um.total; info[2] = um.NF;
  int cnt = 0;
  for (int u = 0; u < um.total; ++u) {
    int J, I0, I1; bool hd;
    um.decode(u, J, I0, I1, hd);
    for (int I = I0; I < I1; ++I) {
      if (cnt >= cap) return -1;
      out[4 * cnt] = u; out[4 * cnt + 1] = I; out[4 * cnt + 2] = J; out[4 * cnt + 3] = 0; ++cnt;
    }
    if (hd) {
      if (cnt >= cap) return -1;
      out[4 * cnt] = u; out[4 * cnt + 1] = J; out[4 * cnt + 2] = J; out[4 * cnt + 3] = 1; ++cnt;
    }
  }
  return cnt;
}

extern "C" void unitmap_desc(int j, int G, int P, int upc, int* out /* j,Tn,C,rcpC,KB,NF,total,R,ndj,nsj */) {
  ColDesc d;
  compute_desc(d, j, G, P, upc);
  out[0] = d.j; out[1] = d.Tn; out[2] = d.C; out[3] = d.rcpC; out[4] = d.KB; out[5] = d.NF; out[6] = d.total;
  out[7] = d.R; out[8] = d.ndj; out[9] = d.nsj;
}

// strip-length table (tile::build_strip_table) for CPU tests
extern "C" void unitmap_strip_table(int G, int P, int Tnmax, int cmax, double ov, unsigned char* out) {
  build_strip_table(G, P, Tnmax, cmax, ov, out);
}
