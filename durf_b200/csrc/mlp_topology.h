// Host-side description of the reference MLP / BoxMLP (obbpose_model.py:294-354, 358-418):
// per-layer shapes and offsets into the fp32 parameter blob declared in include/durf_b200.h.
#pragma once

#include <stdint.h>
#include <vector>

#include "../../include/durf_b200.h"

namespace durf {

struct MlpLayout {
  int n_layers;                 // depth + 4
  std::vector<int> in_dim, out_dim;
  std::vector<int64_t> w_off, b_off;
  int64_t total;

  explicit MlpLayout(const DurfMlpTopology& t) {
    n_layers = t.depth + 4;
    int k = t.in_dim;
    for (int i = 0; i < t.depth; ++i) {
      in_dim.push_back(k);
      out_dim.push_back(t.width);
      k = (i % t.skip == 0 && i > 0) ? t.width + t.in_dim : t.width;
    }
    in_dim.push_back(k); out_dim.push_back(1);                           // density
    in_dim.push_back(k); out_dim.push_back(t.width);                     // bottleneck
    in_dim.push_back(t.width + t.cond_dim); out_dim.push_back(t.cond_width);  // condition layer
    in_dim.push_back(t.cond_width); out_dim.push_back(3);                // rgb
    int64_t off = 0;
    for (int i = 0; i < n_layers; ++i) {
      w_off.push_back(off);
      off += (int64_t)in_dim[i] * out_dim[i];
      b_off.push_back(off);
      off += out_dim[i];
    }
    total = off;
  }
};

static inline bool topology_ok(const DurfMlpTopology* t) {
  return t && t->in_dim >= 1 && t->width >= 1 && t->depth >= 1 && t->depth <= 12 && t->skip >= 1 && t->cond_dim >= 1 &&
         t->cond_width >= 1;
}

}  // namespace durf
