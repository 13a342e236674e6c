/* include/b200tfs.h must be plain C (the boundary a cgo / JNI / ctypes binding sees): compiled with
 * gcc -std=c99 -pedantic, linked against libb200tfs.so, run without a GPU. */
#include <stdio.h>

#include "../../include/b200tfs.h"

int main(void) {
  b200tfs_ctx* ctx = 0;
  int rc;
  printf("abi %d sizes %d %d %d %d\n", b200tfs_abi_version(), (int)sizeof(b200tfs_tensor), (int)sizeof(b200tfs_request),
         (int)sizeof(b200tfs_output), (int)sizeof(b200tfs_model_spec));
  rc = b200tfs_create(0, &ctx);
  printf("create %d %s\n", rc, b200tfs_last_error());
  if (rc == B200TFS_OK) b200tfs_destroy(ctx);
  return 0;
}
