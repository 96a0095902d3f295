// Drop-in replacement for the reference's include/GPUImageProcessing.h
// (signatures at /root/reference/include/GPUImageProcessing.h:4-10).
#ifndef GPU_IMAGE_PROCESSING_H
#define GPU_IMAGE_PROCESSING_H

#include <cstddef>

// ref GPUImageProcessing.h:4-5 / GPUImageProcessing.cu:8-21,72-79 -- dst = (float)src.ch0 where mask == 255
void GPUConvertToFloat(unsigned char *src, size_t srcPitch, float *dst, size_t dstPitch, unsigned char *mask, size_t maskPitch,
	int rows, int cols);
// ref GPUImageProcessing.h:6-8 / GPUImageProcessing.cu:23-49,81-91 -- 2x2 "any scribbled" restriction
void GPUPyrDownAnnotation(unsigned char *prevScribbleImage, size_t prevScribblePitch, unsigned char *prevEditedImage,
	size_t prevEditedPitch, int previousRows, int previousCols, unsigned char *currScribbleImage, size_t currScribblePitch,
	unsigned char *currEditedImage, size_t currEditedPitch, int currentRows, int currentCols);
// ref GPUImageProcessing.h:9-10 / GPUImageProcessing.cu:51-70,93-100 -- square brush
void GPUPaintImage(int x, int y, int scribbleColor, int scribbleRadius, unsigned char *editedImage, size_t editedPitch,
	unsigned char *scribbleImage, size_t scribblePitch, int rows, int cols);

#endif
