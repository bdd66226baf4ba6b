/* ssw_host.h -- small host-side helpers shared by the engine sources. */
#ifndef SSW_HOST_H
#define SSW_HOST_H

#include "ssw_common.cuh"

/* grow-only device buffer */
struct SswDevBuf {
	void* p = nullptr;
	size_t cap = 0;
	int ensure(size_t bytes)
	{
		if (bytes <= cap) return 0;
		if (p) cudaFree(p);
		p = nullptr; cap = 0;
		size_t want = bytes + bytes / 8 + 256;
		if (cudaMalloc(&p, want) != cudaSuccess) {
			fprintf(stderr, "[libssw-b200] device allocation of %zu bytes failed\n", want);
			return -1;
		}
		cap = want;
		return 0;
	}
	void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
	template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

/* CUDA-event stopwatch on one stream */
struct SswTimer {
	cudaEvent_t a = nullptr, b = nullptr;
	void start(cudaStream_t s) { if (!a) { cudaEventCreate(&a); cudaEventCreate(&b); } cudaEventRecord(a, s); }
	float stop(cudaStream_t s) { cudaEventRecord(b, s); cudaEventSynchronize(b); float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; }
	~SswTimer() { if (a) { cudaEventDestroy(a); cudaEventDestroy(b); } }
};

#endif
