"""An INDEPENDENT float64 statement of the Kaldi log-mel filterbank the reference implements
(src/fbank.cc:44-69 ProcessWindow, :103-163 Melbanks, :193-244 power spectrum / floor / log) -- plain
numpy with numpy's own FFT, sharing no code with oracle/ce_oracle.c or the CUDA kernel.  It exists to
pin sizes the reference itself cannot run: PK_FBANK_DIM is a hard `#define 40` and the reference aborts
in its Melbanks constructor at 80 bins (src/fbank.cc:155: the lowest filter spans a single FFT bin).
The formula has one parameter, the number of bins; tests first check this implementation against the
Kaldi golden dump for 40 bins (test/data/fbankmat_en-us-hello.wav.txt) and then use it at 80.

TEST INFRASTRUCTURE ONLY."""
import numpy as np

SAMPLE_RATE, FRAME_LEN, FRAME_SHIFT, FFT_LEN = 16000, 400, 160, 512
LOW_FREQ, HIGH_FREQ, PREEMPH = 20.0, 8000.0, 0.97


def mel(f):
    return 1127.0 * np.log(1.0 + np.asarray(f, np.float64) / 700.0)


def mel_weights(num_mel):
    """[num_mel x 257] triangular weights; only FFT bins 0..255 carry weight (src/fbank.cc:108,136),
    strict inequalities at the triangle's feet (:141)."""
    w = np.zeros((num_mel, FFT_LEN // 2 + 1))
    lo, hi = mel(LOW_FREQ), mel(HIGH_FREQ)
    delta = (hi - lo) / (num_mel + 1)
    m = mel(np.arange(FFT_LEN // 2) * (SAMPLE_RATE / FFT_LEN))
    for b in range(num_mel):
        left, center, right = lo + b * delta, lo + (b + 1) * delta, lo + (b + 2) * delta
        inside = (m > left) & (m < right)
        up = (m - left) / (center - left)
        down = (right - m) / (right - center)
        w[b, :FFT_LEN // 2] = np.where(inside, np.where(m <= center, up, down), 0.0)
    return w


def fbank(pcm, num_mel=40):
    """[T x num_mel] log-mel energies of unscaled int16 samples, snip-edges framing, no dither."""
    pcm = np.asarray(pcm, np.float64)
    T = 0 if pcm.size < FRAME_LEN else 1 + (pcm.size - FRAME_LEN) // FRAME_SHIFT
    if T == 0:
        return np.zeros((0, num_mel))
    idx = np.arange(T)[:, None] * FRAME_SHIFT + np.arange(FRAME_LEN)[None, :]
    x = pcm[idx]
    x = x - x.mean(axis=1, keepdims=True)                     # DC removal
    y = x.copy()
    y[:, 1:] -= PREEMPH * x[:, :-1]                           # x[i] -= 0.97 x[i-1], back to front
    y[:, 0] -= PREEMPH * x[:, 0]                              # x[0] -= 0.97 x[0]       (:61)
    ham = 0.54 - 0.46 * np.cos(2.0 * np.pi * np.arange(FRAME_LEN) / (FRAME_LEN - 1))
    spec = np.fft.rfft(y * ham, n=FFT_LEN, axis=1)
    power = spec.real ** 2 + spec.imag ** 2
    e = power @ mel_weights(num_mel).T
    return np.log(np.maximum(e, np.finfo(np.float32).eps))
