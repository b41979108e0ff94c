// Internal hooks between the translation units of libportal_b200.so (not part of the C ABI).
#pragma once
struct pe_ctx;
// CUDA device ordinal of a context, -1 for a compile-only context.
int pe_internal_device(pe_ctx* ctx);
