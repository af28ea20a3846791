#include "mrl_internal.h"
int mrl_expr_launch_zfwd(mrl_context *, void *, const void *, void *, void *, void *, long long, int) {
  return mrl_fail(MRL_ERR_UNSUPPORTED, "expression-compiled nonlinearity not built yet");
}
