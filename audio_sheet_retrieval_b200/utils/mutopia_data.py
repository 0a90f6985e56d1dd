"""Data selection.  The reference loads MSMD pieces through the `msmd` package
(audio_sheet_retrieval/utils/mutopia_data.py:47-98), which is not available here; what every
caller on the hot path uses is only the *pool protocol* (utils/data_pools.py:203-228):

    pool.shape[0]                          number of pairs
    pool[slice | int ndarray] -> [X1 (n,1,160,200) float32 holding 0..255,
                                  X2 (n,1,92,42)  float32 >= 0]

`SyntheticPairPool` implements that protocol with deterministic, seeded, sheet-like and
spectrogram-like data so run_eval / refine_cca / the benches run end to end.
"""
import numpy as np

SHEET_CONTEXT = 200      # exp_configs/*.yaml
SYSTEM_HEIGHT = 160
SPEC_CONTEXT = 42
SPEC_BINS = 92


class SyntheticPairPool(object):
    def __init__(self, n_pairs, seed=23, sheet_dtype=np.float32):
        self.n_pairs = int(n_pairs)
        self.seed = int(seed)
        self.sheet_dtype = sheet_dtype
        self.shape = [self.n_pairs]

    def _pair(self, i):
        rng = np.random.RandomState((self.seed * 1000003 + int(i)) % (2 ** 31 - 1))
        H, W = SYSTEM_HEIGHT, SHEET_CONTEXT
        sheet = np.full((H, W), 255.0, np.float32)
        y0, gap = rng.randint(20, 50), rng.randint(6, 10)
        for s in range(2):
            for l in range(5):
                y = y0 + s * 70 + l * gap
                if y < H:
                    sheet[y] = rng.randint(0, 60)
        for _ in range(rng.randint(10, 40)):
            y, x = rng.randint(0, H - 8), rng.randint(0, W - 8)
            sheet[y:y + rng.randint(3, 8), x:x + rng.randint(3, 8)] = rng.randint(0, 100)
        for _ in range(rng.randint(4, 16)):
            x, ya = rng.randint(0, W - 1), rng.randint(0, H - 30)
            sheet[ya:ya + rng.randint(10, 30), x] = rng.randint(0, 80)
        spec = np.abs(rng.normal(0, 0.02, (SPEC_BINS, SPEC_CONTEXT))).astype(np.float32)
        t = np.arange(SPEC_CONTEXT)
        for _ in range(rng.randint(3, 10)):
            f0, t0, dur = rng.randint(5, 40), rng.randint(0, SPEC_CONTEXT - 4), rng.randint(4, SPEC_CONTEXT)
            env = np.where(t >= t0, np.exp(-(t - t0) / float(dur)), 0.0) * rng.uniform(0.5, 2.0)
            for hnum in range(1, 6):
                f = int(f0 + 12 * np.log2(hnum))
                if f < SPEC_BINS:
                    spec[f] += (env / hnum).astype(np.float32)
        return sheet, spec

    def __getitem__(self, key):
        if isinstance(key, slice):
            idx = np.arange(*key.indices(self.n_pairs))
        else:
            idx = np.atleast_1d(np.asarray(key)).astype(int)
        X1 = np.empty((len(idx), 1, SYSTEM_HEIGHT, SHEET_CONTEXT), np.float32)
        X2 = np.empty((len(idx), 1, SPEC_BINS, SPEC_CONTEXT), np.float32)
        for j, i in enumerate(idx):
            X1[j, 0], X2[j, 0] = self._pair(i)
        if self.sheet_dtype != np.float32:
            X1 = X1.astype(self.sheet_dtype)
        return [X1, X2]


def load_audio_score_retrieval(split_file=None, config_file=None, test_only=False, n_train=25000, n_valid=1000,
                               n_test=2000, seed=23):
    """-> dict(train, valid, test) of pools (mutopia_data.py:47-98 returns the same keys)."""
    test = SyntheticPairPool(n_test, seed=seed + 2)
    if test_only:
        return dict(train=None, valid=None, test=test)
    return dict(train=SyntheticPairPool(n_train, seed=seed), valid=SyntheticPairPool(n_valid, seed=seed + 1), test=test)


def load_msmd_audio_score_retrieval(split_file=None, config_file=None, test_only=False):
    """The reference's loader (mutopia_data.py:47-98) goes through the `msmd` package and its data set.  Neither
    ships with this repo; when `msmd` is importable the real pools would be built here -- until then this raises
    instead of handing out synthetic data under the real data set's name."""
    try:
        import msmd  # noqa: F401
    except ImportError:
        raise RuntimeError("--data mutopia needs the MSMD data set and the `msmd` package "
                           "(audio_sheet_retrieval/utils/mutopia_data.py:8-13), which are not installed.  "
                           "Use --data synthetic for the seeded synthetic pools (results are tagged 'synthetic').")
    raise NotImplementedError("MSMD loading is out of scope of the hot path (SURVEY.md 2.1 #17): plug the reference's "
                              "load_audio_score_retrieval in here; it must return dict(train, valid, test) of pools")
