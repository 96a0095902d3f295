// Drop-in replacement for the reference's include/GPUDepthEffect.h
// (signatures at /root/reference/include/GPUDepthEffect.h:4-9).
// Three free functions with C++ linkage; the parameter TYPES are the reference's (same mangled
// symbols, tests/test_abi_symbols.py), the rest is this library's documentation.  DEVICE pointers,
// row pitches in BYTES, BGR = u8 x 3 interleaved, depth = fp32 in [0, 255].  Asynchronous on the
// library's stream, like the reference's.  rtdd_effects_fused / rtdd_frame_effects (include/rtdd.h)
// produce all three from one read of the inputs.
#ifndef GPU_DEPTH_EFFECT_H
#define GPU_DEPTH_EFFECT_H

#include <cstddef>

// Refocus: each pixel becomes the mean of the image over a square window whose side grows with its
// depth, a = (int)(K * d / 255), K = (int)(0.025 * image diagonal); empty window -> unchanged.
// ref GPUDepthEffect.h:4-5 / GPUDepthEffect.cu:29-72,105-113.  Here: exact integer summed-area
// table (bit-identical to the reference's tap-by-tap fp32 sums); rtdd_defocus underneath.
void GPUSimulateDefocus(
	unsigned char *imageBgr,
	size_t imageBgrPitch,
	float *depth,
	size_t depthPitch,
	unsigned char *resultBgr,
	size_t resultBgrPitch,
	int rows,
	int cols);

// Desaturation: every channel moves from its colour towards the gray value by f = d / 255.
// ref GPUDepthEffect.h:6-7 / GPUDepthEffect.cu:8-27,95-103; rtdd_desaturate underneath.
void GPUSimulateDesaturation(
	unsigned char *imageBgr,
	size_t imageBgrPitch,
	unsigned char *imageGray,         // u8, may be larger than rows x cols (only the pitch matters)
	size_t imageGrayPitch,
	float *depth,
	size_t depthPitch,
	unsigned char *resultBgr,
	size_t resultBgrPitch,
	int rows,
	int cols);

// Haze: every channel moves towards white with transmission t = expf(-2 d / 255).
// ref GPUDepthEffect.h:8-9 / GPUDepthEffect.cu:74-93,115-123; rtdd_haze underneath.
void GPUSimulateHaze(
	unsigned char *imageBgr,
	size_t imageBgrPitch,
	float *depth,
	size_t depthPitch,
	unsigned char *resultBgr,
	size_t resultBgrPitch,
	int rows,
	int cols);

#endif
