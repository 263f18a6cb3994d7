// Shared between capi.cu and prover.cu: the opaque context type and the exception -> status translation.
#pragma once
#include "../../include/taiga_b200.h"
#include "common.cuh"

struct tb_ctx { tb::Ctx c; };

#define TB_API_BEGIN(ctx) if (!(ctx)) return TB_ERR_INVALID; try {
#define TB_API_END(ctx)                                                                         \
  return TB_OK; }                                                                               \
  catch (const tb::CudaError& e) { (ctx)->c.last_error = e.what(); cudaGetLastError(); return TB_ERR_CUDA; }        \
  catch (const tb::ConstraintError& e) { (ctx)->c.last_error = e.what(); return TB_ERR_CONSTRAINT; }                \
  catch (const std::invalid_argument& e) { (ctx)->c.last_error = e.what(); return TB_ERR_INVALID; }                 \
  catch (const std::exception& e) { (ctx)->c.last_error = e.what(); return TB_ERR_INTERNAL; }
