// Return codes of the C ABI.  Numeric values follow the reference's public enum
// (include/finufft_errors.h:9-44) so existing callers can interpret them unchanged.
#pragma once

namespace b200 {
enum ErrorCode : int {
  OK                       = 0,
  ERR_MAXNALLOC            = 2,   // fine grid larger than 1e12 points
  ERR_SPREAD_BOX_SMALL     = 3,   // a fine-grid dimension is below 2*ns
  ERR_UPSAMPFAC_TOO_SMALL  = 7,   // sigma <= 1
  ERR_HORNER_WRONG_BETA    = 8,   // device API: nonstandard sigma with gpu_kerevalmeth = 1
  ERR_NTRANS_NOTVALID      = 9,
  ERR_TYPE_NOTVALID        = 10,
  ERR_ALLOC                = 11,
  ERR_DIM_NOTVALID         = 12,
  ERR_NDATA_NOTVALID       = 14,  // a size does not fit the 32-bit device index range
  ERR_CUDA_FAILURE         = 15,
  ERR_PLAN_NOTVALID        = 16,
  ERR_METHOD_NOTVALID      = 17,
  ERR_BINSIZE_NOTVALID     = 18,
  ERR_INSUFFICIENT_SHMEM   = 19,
  ERR_NUM_NU_PTS_INVALID   = 20,
  ERR_INVALID_ARGUMENT     = 21,
  ERR_LOCK_FUNS_INVALID    = 22,
  ERR_NTHREADS_NOTVALID    = 23,
  ERR_KERFORMULA_NOTVALID  = 24,
  ERR_UNKNOWN_EXCEPTION    = 25,
  ERR_EPS_TOO_SMALL        = 26,
  ERR_PSWF_SETUP           = 27,
};

struct Failure {  // thrown inside the engine, mapped to an int at the C boundary
  int code;
};
}  // namespace b200
