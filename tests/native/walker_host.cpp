// walker_host.cpp - compiles the decode kernels' tag walker (min-tfs-client_b200/csrc/walker.h) for
// the HOST so its logic can be unit-tested without a GPU (tests/test_walker_host.py replays
// tests/golden/decode.json through it).  Test infrastructure only: the product never loads this.
#include "../../min-tfs-client_b200/csrc/walker.h"

using namespace b200tfs;

extern "C" {

int wh_parse_response(const uint8_t* wire, uint64_t len, int max_outputs, b200tfs_output* outs /* max_outputs + 1 */, int* n_outs,
                      b200tfs_model_spec* spec) {
  if (len > 0x7FFFFFFFull) return B200TFS_E_PARSE;
  Cursor c;
  cur_open_host(c, wire, (uint32_t)len);
  return walk_response(c, max_outputs, outs, n_outs, spec);
}

int wh_parse_tensor(const uint8_t* wire, uint64_t len, b200tfs_output* out) {
  if (len > 0x7FFFFFFFull) return B200TFS_E_PARSE;
  Cursor c;
  cur_open_host(c, wire, (uint32_t)len);
  return walk_tensor_proto(c, out);
}

int wh_sizeof_output(void) { return (int)sizeof(b200tfs_output); }
}
