// Binary catalogue ingest: a NumPy .npy file (2-D, C order,
// little-endian float64 or float32, shape (N, ncols)) goes from the page cache to
// the 32-byte particle records in HBM without a host-side parse:
//   mmap -> pinned staging (a few host threads) -> H2D on the copy stream
//        -> k_assemble: column selection, w = wcomp * wfkp and the catalogue sums
// while the next chunk is being copied.  It produces what read_ascii_data()
// produces for the same numbers (io/read_ascii.c:868-902): records {x, y, z, w},
// sum wcomp, sum w^2, sum wcomp wfkp^2 n(z); absent columns mean wcomp = 1,
// wfkp = 1, n(z) = 0, and for simulation boxes w = wcomp.  The reference reads
// ASCII through libast expressions (minutes for 1e8 lines) and has only a stub
// for anything else (src/read_cata.c:126-131).
//
// The sums are accumulated per block and added on the host in block order, so a
// file always gives the same bits.

#include "psb_internal.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdlib>
#include <cstring>
#include <string>

namespace psb {

// parses the header of a .npy file; returns 0 and the payload offset, the shape and
// the element size (8 / 4), or -1 with the error set
int npy_probe(const char *path, size_t *offset, size_t *nrow, int *ncol, int *elem) {
  FILE *f = fopen(path, "rb");
  if (!f) { set_error("cannot open file for reading: `%s'\n", path); return -1; }
  unsigned char pre[12];
  if (fread(pre, 1, 10, f) != 10 || memcmp(pre, "\x93NUMPY", 6) != 0) {
    fclose(f);
    set_error("not a .npy file: `%s'\n", path);
    return -1;
  }
  size_t hlen = 0, hoff = 10;
  if (pre[6] == 1) hlen = pre[8] | ((size_t) pre[9] << 8);
  else {
    if (fread(pre + 10, 1, 2, f) != 2) { fclose(f); set_error("truncated .npy header: `%s'\n", path); return -1; }
    hlen = pre[8] | ((size_t) pre[9] << 8) | ((size_t) pre[10] << 16) | ((size_t) pre[11] << 24);
    hoff = 12;
  }
  if (hlen > (1u << 20)) { fclose(f); set_error("unreasonable .npy header: `%s'\n", path); return -1; }
  std::string h(hlen, '\0');
  if (fread(&h[0], 1, hlen, f) != hlen) { fclose(f); set_error("truncated .npy header: `%s'\n", path); return -1; }
  fclose(f);
  auto value_of = [&](const char *key) -> std::string {
    const size_t k = h.find(key);
    if (k == std::string::npos) return "";
    size_t p = h.find(':', k);
    if (p == std::string::npos) return "";
    ++p;
    while (p < h.size() && h[p] == ' ') ++p;
    size_t e = p;
    if (h[p] == '(') e = h.find(')', p);
    else if (h[p] == '\'') e = h.find('\'', p + 1);
    else e = h.find_first_of(",}", p);
    if (e == std::string::npos) return "";
    return h.substr(p, e - p + 1);
  };
  const std::string descr = value_of("'descr'"), order = value_of("'fortran_order'"),
      shape = value_of("'shape'");
  if (descr == "'<f8'" || descr == "'|f8'") *elem = 8;
  else if (descr == "'<f4'" || descr == "'|f4'") *elem = 4;
  else { set_error("unsupported .npy dtype %s (need <f8 or <f4): `%s'\n", descr.c_str(), path); return -1; }
  if (order.compare(0, 5, "False") != 0) { set_error("Fortran-ordered .npy not supported: `%s'\n", path); return -1; }
  unsigned long long a = 0, b = 0;
  if (sscanf(shape.c_str(), "(%llu, %llu)", &a, &b) != 2 || b < 1 || b > 64) {
    set_error("the .npy array must be 2-D (N, ncols): `%s' has shape %s\n", path, shape.c_str());
    return -1;
  }
  *offset = hoff + hlen;
  *nrow = (size_t) a;
  *ncol = (int) b;
  return 0;
}

namespace {

struct Cols { int pos[3], wcomp, wfkp, nz, ncol, issim; };

constexpr int ASM_BLOCKS = 148 * 4, ASM_THREADS = 256;

// rows of `raw` -> records + per-block partial sums (sumw, sumw2, sumw2n)
template <typename T>
__global__ void __launch_bounds__(ASM_THREADS) k_assemble(const T *__restrict__ raw, size_t n,
    Cols cl, double2 *__restrict__ rec, double *__restrict__ partial) {
  double sw = 0, sw2 = 0, sw2n = 0;
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n;
       i += (size_t) gridDim.x * blockDim.x) {
    const T *row = raw + i * cl.ncol;
    const double x = (double) row[cl.pos[0]], y = (double) row[cl.pos[1]], z = (double) row[cl.pos[2]];
    const double wc = cl.wcomp >= 0 ? (double) row[cl.wcomp] : 1.0;
    double w = wc;
    if (!cl.issim) {                                        // io/read_ascii.c:883-902
      const double wf = cl.wfkp >= 0 ? (double) row[cl.wfkp] : 1.0;
      const double nz = cl.nz >= 0 ? (double) row[cl.nz] : 0.0;
      w = __dmul_rn(wc, wf);
      sw2 = __dadd_rn(sw2, __dmul_rn(w, w));
      sw2n = __dadd_rn(sw2n, __dmul_rn(__dmul_rn(__dmul_rn(wc, wf), wf), nz));
    }
    sw = __dadd_rn(sw, wc);
    rec[2 * i] = make_double2(x, y);
    rec[2 * i + 1] = make_double2(z, w);
  }
  __shared__ double red[3][ASM_THREADS];
  red[0][threadIdx.x] = sw; red[1][threadIdx.x] = sw2; red[2][threadIdx.x] = sw2n;
  __syncthreads();
  for (int s = ASM_THREADS / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s)
      for (int q = 0; q < 3; q++) red[q][threadIdx.x] += red[q][threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0)
    for (int q = 0; q < 3; q++) partial[3 * blockIdx.x + q] = red[q][0];
}

}  // namespace

int assemble_blocks() { return ASM_BLOCKS; }

int launch_assemble(const void *raw, int elem, size_t n, const int pos[3], int wcomp, int wfkp, int nz,
    int ncol, int issim, double *rec, double *partial, cudaStream_t st) {
  Cols cl;
  for (int a = 0; a < 3; a++) cl.pos[a] = pos[a];
  cl.wcomp = wcomp; cl.wfkp = wfkp; cl.nz = nz; cl.ncol = ncol; cl.issim = issim;
  if (elem == 8)
    k_assemble<double><<<ASM_BLOCKS, ASM_THREADS, 0, st>>>(static_cast<const double *>(raw), n, cl,
        reinterpret_cast<double2 *>(rec), partial);
  else
    k_assemble<float><<<ASM_BLOCKS, ASM_THREADS, 0, st>>>(static_cast<const float *>(raw), n, cl,
        reinterpret_cast<double2 *>(rec), partial);
  PSB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace psb
