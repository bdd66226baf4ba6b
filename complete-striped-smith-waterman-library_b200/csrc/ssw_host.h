/* ssw_host.h -- small host-side helpers shared by the engine sources. */
#ifndef SSW_HOST_H
#define SSW_HOST_H

#include "ssw_common.cuh"

/* counts device allocations and frees made through SswDevBuf: the engine re-reads the free device memory only when
 * this changed (cudaMemGetInfo costs milliseconds while a monitoring tool polls the driver) */
inline unsigned long long& ssw_alloc_epoch() { static unsigned long long n = 1; return n; }

/* grow-only device buffer */
struct SswDevBuf {
	void* p = nullptr;
	size_t cap = 0;
	int ensure(size_t bytes)
	{
		if (bytes <= cap) return 0;
		++ssw_alloc_epoch();
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
	void release() { ++ssw_alloc_epoch(); if (p) cudaFree(p); p = nullptr; cap = 0; }
	template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

/* free memory of the current device, re-read only after one of our buffers was (re)allocated */
inline size_t ssw_free_device_bytes()
{
	static size_t cache[64];
	static unsigned long long epoch[64];
	int dev = 0;
	cudaGetDevice(&dev);
	dev &= 63;
	if (epoch[dev] != ssw_alloc_epoch()) {
		size_t free_b = 0, total_b = 0;
		cudaMemGetInfo(&free_b, &total_b);
		cache[dev] = free_b;
		epoch[dev] = ssw_alloc_epoch();
	}
	return cache[dev];
}

/* CUDA-event stopwatch on one stream */
struct SswTimer {
	cudaEvent_t a = nullptr, b = nullptr;
	void start(cudaStream_t s) { if (!a) { cudaEventCreate(&a); cudaEventCreate(&b); } cudaEventRecord(a, s); }
	float stop(cudaStream_t s) { cudaEventRecord(b, s); cudaEventSynchronize(b); float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; }
	~SswTimer() { if (a) { cudaEventDestroy(a); cudaEventDestroy(b); } }
};

#endif
