// TEMPORARY during bring-up: entry points not implemented yet report an error (never a fallback).
#include "common.cuh"
using namespace rpcc;
#define NOT_YET(name) do { set_error(name ": not implemented yet"); return RPCC_ERR_ARG; } while (0)
extern "C" int rpcc_ground_fit_batch(const float*, const float*, int, int, int, uint64_t, float*, void*, size_t, void*) { NOT_YET("rpcc_ground_fit_batch"); }
extern "C" size_t rpcc_ground_fit_workspace(int, int, int) { return 0; }
extern "C" size_t rpcc_decode_workspace(int, int, int, int) { return 0; }
extern "C" void* rpcc_encoder_stream(rpcc_encoder*) { return nullptr; }
extern "C" void* rpcc_encoder_device_buffer(rpcc_encoder*, const char*) { return nullptr; }
