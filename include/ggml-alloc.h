/* Intentionally empty: the reference bark.h includes "ggml-alloc.h" (bark.h:20-22) only for types that never
 * cross the bark.h boundary.  This shim keeps unchanged callers compiling; the EnCodec decoder and the
 * tensor runtime are internal to libbark_b200 (hand-written CUDA), not a re-exported ggml/encodec API. */
#pragma once
