// Drop-in replacement for the reference's include/GPUImageProcessing.h
// (signatures at /root/reference/include/GPUImageProcessing.h:4-10).
// Three free functions with C++ linkage; the parameter TYPES are the reference's (the mangled
// symbols must match for main.cpp to link unchanged -- tests/test_abi_symbols.py), everything
// else here is this library's documentation of them.  All image pointers are DEVICE pointers,
// every pitch is a row pitch in BYTES.  The functions enqueue on the library's stream and return
// without a device sync, like the reference's.
#ifndef GPU_IMAGE_PROCESSING_H
#define GPU_IMAGE_PROCESSING_H

#include <cstddef>

// Dirichlet injection: depth = (float)colour channel 0 wherever the annotation mask is 255.
// ref GPUImageProcessing.h:4-5 / GPUImageProcessing.cu:8-21,72-79; rtdd_convert_to_float underneath.
void GPUConvertToFloat(
	unsigned char *annotatedBgr,      // u8 x 3 interleaved, only channel 0 is read
	size_t annotatedBgrPitch,
	float *depth,                     // fp32 plane, written only under the mask
	size_t depthPitch,
	unsigned char *annotationMask,    // u8, 255 = scribbled
	size_t annotationMaskPitch,
	int rows,
	int cols);

// Annotation restriction to the next coarser level: output pixel (x, y) looks at input rows
// {2y-1, 2y} x columns {2x-1, 2x}; the LAST scribbled one in row-major order wins; nothing is
// written where none is scribbled (coarse planes keep older strokes).
// ref GPUImageProcessing.h:6-8 / GPUImageProcessing.cu:23-49,81-91; rtdd_pyrdown_annotation underneath.
void GPUPyrDownAnnotation(
	unsigned char *fineMask,          // u8, 255 = scribbled
	size_t fineMaskPitch,
	unsigned char *fineBgr,           // u8 x 3, channel 0 carries the scribble depth
	size_t fineBgrPitch,
	int fineRows,
	int fineCols,
	unsigned char *coarseMask,        // outputs, updated in place
	size_t coarseMaskPitch,
	unsigned char *coarseBgr,
	size_t coarseBgrPitch,
	int coarseRows,
	int coarseCols);

// Square brush of side 2 * (radius / 2) + 1 centred on (x, y), clipped to the image: the three
// colour channels take `depthValue`, the mask takes 255.
// ref GPUImageProcessing.h:9-10 / GPUImageProcessing.cu:51-70,93-100; rtdd_paint underneath.
void GPUPaintImage(
	int x,
	int y,
	int depthValue,
	int radius,
	unsigned char *annotatedBgr,
	size_t annotatedBgrPitch,
	unsigned char *annotationMask,
	size_t annotationMaskPitch,
	int rows,
	int cols);

#endif
