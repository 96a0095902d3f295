// Drop-in replacement for the reference's include/GPUDepthEffect.h
// (signatures at /root/reference/include/GPUDepthEffect.h:4-9).
#ifndef GPU_DEPTH_EFFECT_H
#define GPU_DEPTH_EFFECT_H

#include <cstddef>

// ref GPUDepthEffect.h:4-5 / GPUDepthEffect.cu:29-72,105-113 -- depth-sized box blur
void GPUSimulateDefocus(unsigned char *originalImage, size_t originalPitch, float *depthImage, size_t depthPitch,
	unsigned char *artisticImage, size_t artisticPitch, int rows, int cols);
// ref GPUDepthEffect.h:6-7 / GPUDepthEffect.cu:8-27,95-103 -- lerp colour -> gray by depth
void GPUSimulateDesaturation(unsigned char *originalImage, size_t originalPitch, unsigned char *grayImage, size_t grayPitch,
	float *depthImage, size_t depthPitch, unsigned char *artisticImage, size_t artisticPitch, int rows, int cols);
// ref GPUDepthEffect.h:8-9 / GPUDepthEffect.cu:74-93,115-123 -- lerp colour -> white by exp(-2 d/255)
void GPUSimulateHaze(unsigned char *originalImage, size_t originalPitch, float *depthImage, size_t depthPitch,
	unsigned char *artisticImage, size_t artisticPitch, int rows, int cols);

#endif
