// Internal hooks between the translation units of libportal_b200.so (not part of the C ABI).
#pragma once
struct pe_ctx;
// CUDA device ordinal of a context, -1 for a compile-only context.
int pe_internal_device(pe_ctx* ctx);
struct pe_target;
// pe_render / pe_render_rgba8 through program instance 0 or 1 (each has its own uniform block on the device): frames rendered
// through different instances on different streams may overlap.
int pe_internal_render(pe_ctx* ctx, const pe_target* target, void* out_device, void* stream, int rgba8, int instance);
