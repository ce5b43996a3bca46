// featurize.cu — the reference's feature extraction on the GPU: decoded 16-bit PCM -> the [B, T, 80]
// float32 `sequences` that inference_fn consumes.  Replaces load_sample's arithmetic
// (asr/input_functions.py:156-262): python_speech_features' mfcc / logfbank / delta as the reference
// parameterises them (asr/input_functions.py:264-318: 25 ms / 10 ms rectangular frames, pre-emphasis
// 0.97, nfft 1024, 80 triangular mel filters 64 Hz .. Nyquist, log; MFCC: orthonormal DCT-II, first 40
// coefficients, lifter 22, c0 := log frame energy, + 40 delta features over +-2 frames) and
// __feature_normalization (:321-349).  File I/O (wavfile.read) stays on the host.
//
//   frame_features_kernel  one CTA per frame (grid-stride): pre-emphasis while loading, zero-padded
//                          1024-point radix-2 FFT in shared memory (fp32, twiddles from sincospi),
//                          power spectrum, frame energy, mel filters, log, (DCT + lifter)
//   finalize_kernel        one CTA per utterance: delta features, optional frame dropping,
//                          per-utterance mean / standard deviation (accumulated in fp64), zero padding
// Both are latency/HBM-trivial next to the model (32 x 10 s: 32k frames, ~2 GFLOP).
#include "common.cuh"

#include <math.h>

namespace ctcasr {
namespace feat {

constexpr int NFFT = 1024, LOGN = 10, NBIN = NFFT / 2 + 1;
constexpr int MAXF = 128;                       // filters / features
constexpr float PREEMPH = 0.97f;                // asr/input_functions.py:291, :317
constexpr int CEPLIFTER = 22;                   // :291
constexpr int DELTA_N = 2;                      // :293

struct Params {
    const int16_t *audio; int B, max_samples; const int *num_samples;
    int frame_len, frame_step, nfilt, numcep, type, norm, drop;
    float *raw; int Tfull;                      // [B][Tfull][nfilt]
    float *out; int Tmax; int *num_frames;      // [B][Tmax][nfilt]
    int bins[MAXF + 2];                         // filterbank edges (psf.get_filterbanks)
};

__device__ __forceinline__ int frames_of(int n, int len, int step) { return n <= len ? 1 : 1 + (n - len + step - 1) / step; }

__global__ void __launch_bounds__(256) frame_features_kernel(const Params p)
{
    __shared__ float re[NFFT], im[NFFT];
    __shared__ float twr[NFFT / 2], twi[NFFT / 2];
    __shared__ float pw[NBIN + 3];
    __shared__ float lfb[MAXF];
    __shared__ float red[8];
    const int b = blockIdx.y, tid = threadIdx.x;
    const int N = p.num_samples[b];
    const int Tb = frames_of(N, p.frame_len, p.frame_step);
    const int16_t *a = p.audio + (size_t)b * p.max_samples;
    for (int k = tid; k < NFFT / 2; k += 256) {
        float s, c;
        sincospif(-2.0f * (float)k / (float)NFFT, &s, &c);        // exp(-2 pi i k / N)
        twr[k] = c; twi[k] = s;
    }
    for (int t = blockIdx.x; t < Tb; t += gridDim.x) {
        __syncthreads();
        // sigproc.preemphasis + framesig (zero padding past the signal), stored bit-reversed
        for (int i = tid; i < NFFT; i += 256) {
            float v = 0.f;
            const int idx = t * p.frame_step + i;
            if (i < p.frame_len && idx < N)
                v = idx > 0 ? (float)a[idx] - PREEMPH * (float)a[idx - 1] : (float)a[0];
            const int r = (int)(__brev((unsigned)i) >> (32 - LOGN));
            re[r] = v; im[r] = 0.f;
        }
        __syncthreads();
        for (int s = 1; s <= LOGN; ++s) {
            const int half = 1 << (s - 1);
            for (int j = tid; j < NFFT / 2; j += 256) {
                const int pos = j & (half - 1), i0 = ((j >> (s - 1)) << s) + pos, i1 = i0 + half;
                const int tw = pos << (LOGN - s);
                const float wr = twr[tw], wi = twi[tw];
                const float xr = re[i1], xi = im[i1];
                const float tr = wr * xr - wi * xi, ti = wr * xi + wi * xr;
                const float ur = re[i0], ui = im[i0];
                re[i0] = ur + tr; im[i0] = ui + ti;
                re[i1] = ur - tr; im[i1] = ui - ti;
            }
            __syncthreads();
        }
        // sigproc.powspec and the frame energy (fixed-order tree: deterministic)
        float e = 0.f;
        for (int k = tid; k < NBIN; k += 256) {
            const float v = (re[k] * re[k] + im[k] * im[k]) * (1.0f / NFFT);
            pw[k] = v;
            e += v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
        if ((tid & 31) == 0) red[tid >> 5] = e;
        __syncthreads();
        float energy = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) energy += red[w];
        if (energy == 0.f) energy = 2.220446049250313e-16f;       // numpy.finfo(float).eps, as psf
        // triangular filters (psf.get_filterbanks) and log
        if (tid < p.nfilt) {
            const int b0 = p.bins[tid], b1 = p.bins[tid + 1], b2 = p.bins[tid + 2];
            float acc = 0.f;
            for (int i = b0; i < b1; ++i) acc += pw[i] * ((float)(i - b0) / (float)(b1 - b0));
            for (int i = b1; i < b2; ++i) acc += pw[i] * ((float)(b2 - i) / (float)(b2 - b1));
            if (acc == 0.f) acc = 2.220446049250313e-16f;
            lfb[tid] = logf(acc);
        }
        __syncthreads();
        float *row = p.raw + ((size_t)b * p.Tfull + t) * p.nfilt;
        if (p.type == 0) {                      // 'mel': psf.logfbank
            if (tid < p.nfilt) row[tid] = lfb[tid];
        } else if (tid < p.numcep) {            // 'mfcc': dct(type 2, ortho)[:numcep], lifter, c0 := log(energy)
            const int Nf = p.nfilt;
            float acc = 0.f;
            for (int n = 0; n < Nf; ++n) acc += lfb[n] * cospif((float)(tid * (2 * n + 1)) / (float)(2 * Nf));
            acc *= tid == 0 ? rsqrtf((float)Nf) : sqrtf(2.0f / (float)Nf);
            acc *= 1.0f + (CEPLIFTER / 2.0f) * sinpif((float)tid / (float)CEPLIFTER);
            row[tid] = tid == 0 ? logf(energy) : acc;
        }
    }
}

// value of feature j at ORIGINAL frame t: cepstra / filterbank straight from raw, deltas by psf.delta
__device__ __forceinline__ float feature_at(const Params &p, const float *raw_b, int Tb, int t, int j)
{
    if (p.type == 0 || j < p.numcep) return raw_b[(size_t)t * p.nfilt + j];
    const int c = j - p.numcep;
    float acc = 0.f;
#pragma unroll
    for (int n = -DELTA_N; n <= DELTA_N; ++n) {
        const int tt = min(max(t + n, 0), Tb - 1);                // numpy.pad(mode='edge')
        acc += (float)n * raw_b[(size_t)tt * p.nfilt + c];
    }
    return acc / 10.0f;                                            // 2 * (1 + 4)
}

__global__ void __launch_bounds__(128) finalize_kernel(const Params p)
{
    __shared__ double sh[2][4];
    const int b = blockIdx.x, j = threadIdx.x;
    const int Tb = frames_of(p.num_samples[b], p.frame_len, p.frame_step);
    const int Tk = p.drop ? (Tb + 1) / 2 : Tb;                    // sample[::2]
    const int stride = p.drop ? 2 : 1;
    const float *raw_b = p.raw + (size_t)b * p.Tfull * p.nfilt;
    float *out_b = p.out + (size_t)b * p.Tmax * p.nfilt;
    const bool live = j < p.nfilt;
    double mean = 0.0, inv_std = 1.0;
    if (p.norm != 0) {
        double s = 0.0;
        if (live) for (int t = 0; t < Tk; ++t) s += (double)feature_at(p, raw_b, Tb, t * stride, j);
        if (p.norm == 2) {                      // 'local_scalar': one mean / std for the whole sample
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if ((j & 31) == 0) sh[0][j >> 5] = s;
            __syncthreads();
            s = sh[0][0] + sh[0][1] + sh[0][2] + sh[0][3];
            mean = s / ((double)Tk * p.nfilt);
        } else {
            mean = s / (double)Tk;
        }
        double v = 0.0;
        if (live) for (int t = 0; t < Tk; ++t) { const double d = (double)feature_at(p, raw_b, Tb, t * stride, j) - mean; v += d * d; }
        if (p.norm == 2) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((j & 31) == 0) sh[1][j >> 5] = v;
            __syncthreads();
            v = sh[1][0] + sh[1][1] + sh[1][2] + sh[1][3];
            inv_std = 1.0 / sqrt(v / ((double)Tk * p.nfilt));
        } else {
            inv_std = 1.0 / sqrt(v / (double)Tk);       // numpy.std: population; 0 -> inf/nan like the reference
        }
    }
    if (live) {
        for (int t = 0; t < Tk && t < p.Tmax; ++t)
            out_b[(size_t)t * p.nfilt + j] = (float)(((double)feature_at(p, raw_b, Tb, t * stride, j) - mean) * inv_std);
        for (int t = Tk; t < p.Tmax; ++t) out_b[(size_t)t * p.nfilt + j] = 0.f;     // padded_batch fill (asr/input_functions.py:96)
    }
    if (j == 0) p.num_frames[b] = Tk;
}

static int round_half_up(double x) { return (int)floor(x + 0.5); }
static double hz2mel(double hz) { return 2595.0 * log10(1.0 + hz / 700.0); }
static double mel2hz(double mel) { return 700.0 * (pow(10.0, mel / 2595.0) - 1.0); }

}  // namespace feat
}  // namespace ctcasr

using namespace ctcasr;

extern "C" int ctcasr_feature_frames(int num_samples, int sampling_rate)
{
    const int len = feat::round_half_up(0.025 * sampling_rate), step = feat::round_half_up(0.010 * sampling_rate);
    if (num_samples <= 0 || step <= 0) return 0;
    return num_samples <= len ? 1 : 1 + (num_samples - len + step - 1) / step;
}

// psf.get_filterbanks: nfilt + 2 points equally spaced in mel between 64 Hz (f_min, asr/input_functions.py:222)
// and Nyquist, as FFT-bin indices.  Host-only helper (also what the kernels use).
extern "C" int ctcasr_feature_filterbank_bins(int sampling_rate, int num_filters, int32_t *bins)
{
    CTCASR_REQUIRE(bins && sampling_rate > 128 && num_filters >= 1 && num_filters <= feat::MAXF, "filterbank_bins: bad args");
    const double lowmel = feat::hz2mel(64.0), highmel = feat::hz2mel(sampling_rate / 2.0);
    const double step = (highmel - lowmel) / (num_filters + 1);            // numpy.linspace: start + i * step
    for (int i = 0; i < num_filters + 2; ++i) {
        const double mel = i == num_filters + 1 ? highmel : lowmel + i * step;
        bins[i] = (int32_t)floor((feat::NFFT + 1) * feat::mel2hz(mel) / sampling_rate);
    }
    return CTCASR_OK;
}

extern "C" size_t ctcasr_featurize_workspace_bytes(int B, int max_samples, int sampling_rate, int num_features)
{
    if (B < 1 || max_samples < 1 || num_features < 2 || num_features > feat::MAXF) return 0;
    const size_t T = (size_t)ctcasr_feature_frames(max_samples, sampling_rate);
    return align_up((size_t)B * sizeof(int), 256) + align_up((size_t)B * T * num_features * sizeof(float), 256);
}

extern "C" int ctcasr_featurize(const int16_t *audio, int B, int max_samples, const int32_t *num_samples_host,
                                int feature_type, int normalization, int drop_every_second_frame,
                                int sampling_rate, int num_features,
                                float *features, int Tmax, int32_t *num_frames, void *ws, size_t ws_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    CTCASR_REQUIRE(audio && num_samples_host && features && num_frames && B >= 1 && max_samples >= 1, "featurize: bad args");
    CTCASR_REQUIRE(feature_type == 0 || feature_type == 1, "Requested feature type of %d isn't supported.", feature_type);          // asr/input_functions.py:196
    CTCASR_REQUIRE(normalization >= 0 && normalization <= 2, "Requested feature normalization method %d is invalid.", normalization);  // :200
    CTCASR_REQUIRE(num_features >= 2 && num_features <= feat::MAXF, "featurize: num_features %d", num_features);
    CTCASR_REQUIRE(feature_type == 0 || num_features % 2 == 0, "num_features is not a multiple of 2.");                             // :283
    feat::Params p;
    p.frame_len = feat::round_half_up(0.025 * sampling_rate);
    p.frame_step = feat::round_half_up(0.010 * sampling_rate);
    CTCASR_REQUIRE(p.frame_len >= 1 && p.frame_len <= feat::NFFT && p.frame_step >= 1, "featurize: sampling rate %d", sampling_rate);
    int longest = 0;
    for (int b = 0; b < B; ++b) {
        CTCASR_REQUIRE(num_samples_host[b] >= 401 && num_samples_host[b] <= max_samples,
                       "Sample length %d to short (or longer than the buffer): utterance %d", num_samples_host[b], b);           // :214
        longest = num_samples_host[b] > longest ? num_samples_host[b] : longest;
    }
    p.Tfull = ctcasr_feature_frames(max_samples, sampling_rate);
    const int Tl = ctcasr_feature_frames(longest, sampling_rate);
    const int Tneed = drop_every_second_frame ? (Tl + 1) / 2 : Tl;
    CTCASR_REQUIRE(Tmax >= Tneed, "featurize: output holds %d frames, the longest utterance has %d", Tmax, Tneed);
    const size_t need = ctcasr_featurize_workspace_bytes(B, max_samples, sampling_rate, num_features);
    if (!ws || ws_bytes < need) return fail(CTCASR_ERR_WORKSPACE, "featurize: workspace %zu < %zu", ws_bytes, need);
    int *d_len = reinterpret_cast<int *>(ws);
    p.raw = reinterpret_cast<float *>(reinterpret_cast<char *>(ws) + align_up((size_t)B * sizeof(int), 256));
    CTCASR_CUDA_CHECK(cudaMemcpyAsync(d_len, num_samples_host, (size_t)B * sizeof(int), cudaMemcpyHostToDevice, stream));
    p.audio = audio; p.B = B; p.max_samples = max_samples; p.num_samples = d_len;
    p.nfilt = num_features; p.numcep = num_features / 2; p.type = feature_type; p.norm = normalization;
    p.drop = drop_every_second_frame ? 1 : 0;
    p.out = features; p.Tmax = Tmax; p.num_frames = num_frames;
    if (int rc = ctcasr_feature_filterbank_bins(sampling_rate, num_features, p.bins)) return rc;
    dim3 grid(Tl < 148 * 4 ? Tl : 148 * 4, B);
    feat::frame_features_kernel<<<grid, 256, 0, stream>>>(p);
    CTCASR_LAUNCH_CHECK();
    feat::finalize_kernel<<<B, 128, 0, stream>>>(p);
    CTCASR_LAUNCH_CHECK();
    return CTCASR_OK;
}
