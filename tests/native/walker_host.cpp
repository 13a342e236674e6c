// walker_host.cpp - compiles the decode kernels' tag walker (min-tfs-client_b200/csrc/walker.h) for
// the HOST so its logic can be unit-tested without a GPU (tests/test_walker_host.py replays
// tests/golden/decode.json through it).  Test infrastructure only: the product never loads this.
#include "../../min-tfs-client_b200/csrc/walker.h"

using namespace b200tfs;

extern "C" {

// spill: room for spill_cap entries (32 bytes each); *spill_used receives how many the record wanted
int wh_parse_response(const uint8_t* wire, uint64_t len, int max_outputs, b200tfs_output* outs /* max_outputs + 1 */, int* n_outs,
                      b200tfs_model_spec* spec, void* spill, uint32_t spill_cap, uint32_t* spill_used) {
  if (len > 0x7FFFFFFFull) return B200TFS_E_PARSE;
  Cursor c;
  cur_open_host(c, wire, (uint32_t)len);
  SpillArea sp{(SpillEntry*)spill, spill ? spill_cap : 0u, 0u};
  const int st = walk_response(c, max_outputs, outs, n_outs, spec, sp);
  if (spill_used) *spill_used = sp.used;
  return st;
}

int wh_parse_tensor(const uint8_t* wire, uint64_t len, b200tfs_output* out, void* spill, uint32_t spill_cap, uint32_t* spill_used) {
  if (len > 0x7FFFFFFFull) return B200TFS_E_PARSE;
  Cursor c;
  cur_open_host(c, wire, (uint32_t)len);
  SpillArea sp{(SpillEntry*)spill, spill ? spill_cap : 0u, 0u};
  const int st = walk_tensor_proto(c, out, sp);
  if (spill_used) *spill_used = sp.used;
  return st;
}

int wh_sizeof_output(void) { return (int)sizeof(b200tfs_output); }
int wh_sizeof_spill_entry(void) { return (int)sizeof(SpillEntry); }
}
