"""`lib.FluidNetDataset` as far as the per-timestep path needs it (reference: pytorch/lib/dataset_load.py:10-108).

rayleighTaylor.py:105-107 builds a test-set instance only to split the training YAML into (conf, mconf) with
`createConfDict()`.  Training data (Mantaflow scenes) and its loader are outside the B200 hot path
(SURVEY.md section 2.1): this class keeps the configuration split and tolerates a missing dataset directory;
indexing it raises."""


class FluidNetDataset:
    def __init__(self, conf, prefix, save_dt, preprocess=False, resume=False, pr_n_threads=0):
        self.conf = dict(conf)
        self.mconf = self.conf.pop('modelParam')
        self.prefix = prefix
        self.save_dt = save_dt
        self.data_dir = self.conf.get('dataDir')
        self.dataset = self.conf.get('dataset')

    def createConfDict(self):
        return self.conf, self.mconf

    def __len__(self):
        return 0

    def __getitem__(self, idx):
        raise NotImplementedError("FluidNetDataset: loading Mantaflow training scenes is outside the B200 "
                                  "per-timestep path (use lib.load_manta_data.loadMantaFile for single states)")
