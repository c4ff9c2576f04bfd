// Host-side helper of the patch plan (jax_fem_b200/patch_plan.py): split the cells of every patch into chunks
// for the fused assembly kernel (fused.cu).
//
// The kernel adds the row blocks of a chunk into shared-memory rows in `rounds`; two corners that hit the same
// owned node in the same chunk must be in different rounds.  A greedy first-fit pass per patch (cells in
// ascending order) builds chunks of at most `chunk` cells in which no owned node occurs more than `rmax` times,
// so a chunk needs at most `rmax` rounds whatever the mesh; the round of a corner is the number of earlier cells of
// the chunk that contain the same owned node.  Sequential per patch, a few hundred cell visits each: this is
// plan construction (once per Problem), not the hot path, and runs on the host.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../include/fem_b200.h"

extern "C" int fem_patch_chunks_host(int64_t n_patches, const int64_t* cell_ptr_host, const uint8_t* owned_idx_host,
                                     int nodes_per_cell, int max_owned, int chunk, int rmax,
                                     int32_t* cell_chunk_host, uint8_t* rank_host, int32_t* n_chunks_host) {
  if (n_patches < 0 || !cell_ptr_host || !owned_idx_host || !cell_chunk_host || !rank_host || !n_chunks_host ||
      nodes_per_cell <= 0 || max_owned <= 0 || max_owned > 255 || chunk <= 0 || rmax <= 0 || rmax > 255)
    return FEM_EINVAL;
  std::vector<uint8_t> cnt;     // [chunk][max_owned] occurrences of every owned node
  std::vector<int> size;        // cells in every chunk
  for (int64_t p = 0; p < n_patches; ++p) {
    cnt.clear();
    size.clear();
    int first_open = 0;         // chunks before this one are full
    for (int64_t m = cell_ptr_host[p]; m < cell_ptr_host[p + 1]; ++m) {
      const uint8_t* own = owned_idx_host + m * nodes_per_cell;
      int k = first_open;
      for (;; ++k) {
        if (k == (int)size.size()) {
          size.push_back(0);
          cnt.resize((size_t)(k + 1) * max_owned, 0);
          break;
        }
        if (size[k] >= chunk) continue;
        bool ok = true;
        for (int a = 0; a < nodes_per_cell && ok; ++a)
          if (own[a] != 255 && cnt[(size_t)k * max_owned + own[a]] >= rmax) ok = false;
        if (ok) break;
      }
      for (int a = 0; a < nodes_per_cell; ++a) {
        if (own[a] == 255) {
          rank_host[m * nodes_per_cell + a] = 255;
        } else {
          uint8_t& c = cnt[(size_t)k * max_owned + own[a]];
          rank_host[m * nodes_per_cell + a] = c;     // a cell never holds the same node twice
          ++c;
        }
      }
      cell_chunk_host[m] = k;
      if (++size[k] >= chunk)
        while (first_open < (int)size.size() && size[first_open] >= chunk) ++first_open;
    }
    n_chunks_host[p] = (int32_t)size.size();
  }
  return FEM_OK;
}
