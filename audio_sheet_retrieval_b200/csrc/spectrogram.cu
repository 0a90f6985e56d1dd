// Log-frequency spectrogram front-end (SURVEY 8f "next" row 4, second half): what the reference computes with madmom
// before anything reaches the encoders (tutorials/Embedding Tutorial.ipynb cell 28; audio_sheet_server.py:44-60 streams
// the same chain from the microphone):
//   frames of frame_size samples starting at int(i * sample_rate / fps)  (FramedSignalProcessor, origin='future')
//   -> Hann window -> |FFT| of the first frame_size / 2 bins -> filterbank (bins x bands) -> log10(1 + x)
// One CTA per frame: the windowed frame goes to shared memory in bit-reversed order, a radix-2 FFT runs in place
// (twiddles from a table computed in double precision at first use), magnitudes overwrite the real parts, and one
// thread per band walks the filterbank column (only its non-zero bin range, found on the host).
// Latency-bound and tiny next to the encoders: a 3-minute recording is 3 600 frames.
#include <cmath>
#include <vector>

#include "common.cuh"

namespace asr {

constexpr int SP_THREADS = 256;
constexpr int SP_MAX_LOG2 = 12;              // frame_size <= 4096

__global__ void __launch_bounds__(SP_THREADS)
log_spectrogram_kernel(const float *__restrict__ audio, long long n_samples, int frame_size, int log2n, double hop,
                       const float2 *__restrict__ twiddle, const float *__restrict__ window, const float *__restrict__ fb,
                       const int *__restrict__ band_lo, const int *__restrict__ band_hi, int n_bands, long long n_frames,
                       float *__restrict__ out) {
    extern __shared__ float2 x[];            // frame_size complex values
    const long long frame = blockIdx.x;
    const long long start = (long long)((double)frame * hop);          // int(i * hop_size), as the reference
    for (int i = threadIdx.x; i < frame_size; i += SP_THREADS) {
        const long long s = start + i;
        const float v = s < n_samples ? audio[s] * window[i] : 0.0f;
        x[__brev((unsigned)i) >> (32 - log2n)] = make_float2(v, 0.0f);
    }
    __syncthreads();
    for (int st = 1; st <= log2n; ++st) {
        const int half = 1 << (st - 1);
        for (int b = threadIdx.x; b < frame_size / 2; b += SP_THREADS) {
            const int grp = b >> (st - 1), pos = b & (half - 1);
            const int i0 = (grp << st) + pos, i1 = i0 + half;
            const float2 w = twiddle[pos << (log2n - st)];              // exp(-2 pi i pos / 2^st)
            const float2 a = x[i0], c = x[i1];
            const float2 t = make_float2(c.x * w.x - c.y * w.y, c.x * w.y + c.y * w.x);
            x[i0] = make_float2(a.x + t.x, a.y + t.y);
            x[i1] = make_float2(a.x - t.x, a.y - t.y);
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < frame_size / 2; i += SP_THREADS) {
        const float2 v = x[i];
        x[i].x = sqrtf(v.x * v.x + v.y * v.y);
    }
    __syncthreads();
    for (int band = threadIdx.x; band < n_bands; band += SP_THREADS) {
        float acc = 0.0f;
        for (int bin = band_lo[band]; bin < band_hi[band]; ++bin) acc = fmaf(x[bin].x, fb[(size_t)bin * n_bands + band], acc);
        out[(size_t)band * n_frames + frame] = log10f(acc + 1.0f);
    }
}

struct SpecTables {
    int frame_size = 0;
    float2 *twiddle = nullptr;
    float *window = nullptr;
};
static SpecTables g_spec[ASR_MAX_DEVICES];

}  // namespace asr

using namespace asr;

extern "C" {

int asr_spectrogram_num_frames(int64_t n_samples, int sample_rate, double fps) {
    if (n_samples <= 0 || sample_rate <= 0 || fps <= 0) return 0;
    return (int)ceil((double)n_samples / ((double)sample_rate / fps));
}

int asr_log_spectrogram(const float *audio_dev, int64_t n_samples, int sample_rate, int frame_size, double fps,
                        const float *filterbank_dev, const int32_t *band_lo_dev, const int32_t *band_hi_dev, int n_bands,
                        float *out_dev, void *stream) {
    int rc = ensure_device();
    if (rc) return rc;
    ASR_CHECK_ARG(audio_dev && filterbank_dev && band_lo_dev && band_hi_dev && out_dev, "NULL buffer");
    ASR_CHECK_ARG(n_samples > 0 && sample_rate > 0 && fps > 0 && n_bands >= 1, "bad argument");
    int log2n = 0;
    while ((1 << log2n) < frame_size) ++log2n;
    ASR_CHECK_ARG((1 << log2n) == frame_size && log2n >= 4 && log2n <= SP_MAX_LOG2, "frame_size must be a power of two in [16, 4096]");
    cudaStream_t st = (cudaStream_t)stream;
    const int dev = std::max(0, std::min(current_device(), ASR_MAX_DEVICES - 1));
    SpecTables &t = g_spec[dev];
    if (t.frame_size != frame_size) {        // tables in double precision, once per frame size and device
        std::vector<float2> tw(frame_size / 2);
        std::vector<float> win(frame_size);
        const double pi = 3.14159265358979323846;
        for (int i = 0; i < frame_size / 2; ++i)
            tw[i] = make_float2((float)cos(-2.0 * pi * i / frame_size), (float)sin(-2.0 * pi * i / frame_size));
        for (int i = 0; i < frame_size; ++i) win[i] = (float)(0.5 - 0.5 * cos(2.0 * pi * i / (frame_size - 1)));   // np.hanning
        ASR_CUDA(cudaStreamSynchronize(st));
        cudaFree(t.twiddle); cudaFree(t.window);
        t.twiddle = nullptr; t.window = nullptr; t.frame_size = 0;
        ASR_CUDA(cudaMalloc(&t.twiddle, tw.size() * sizeof(float2)));
        ASR_CUDA(cudaMalloc(&t.window, win.size() * sizeof(float)));
        ASR_CUDA(cudaMemcpy(t.twiddle, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
        ASR_CUDA(cudaMemcpy(t.window, win.data(), win.size() * sizeof(float), cudaMemcpyHostToDevice));
        t.frame_size = frame_size;
    }
    const long long n_frames = asr_spectrogram_num_frames(n_samples, sample_rate, fps);
    log_spectrogram_kernel<<<(unsigned)n_frames, SP_THREADS, (size_t)frame_size * sizeof(float2), st>>>(
        audio_dev, n_samples, frame_size, log2n, (double)sample_rate / fps, t.twiddle, t.window, filterbank_dev, band_lo_dev,
        band_hi_dev, n_bands, n_frames, out_dev);
    ASR_LAUNCH_CHECK();
    return ASR_OK;
}

}  // extern "C"
