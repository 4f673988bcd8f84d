"""Reference module name ``audio_dataset``: the batch composition (FileBatchSampler, audio_dataset.py:202-263) is on
the hot path and implemented; decoding audio files from disk (AudioDataset, :16-199) is out of scope --
``SyntheticAudioDataset`` exposes the same three methods the trainer uses."""
import _bootstrap  # noqa: F401
from cpc_b200.sampler import FileBatchSampler, SyntheticAudioDataset                                 # noqa: F401


class AudioDataset:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("decoding audio files is out of scope of the B200 hot path; pass any dataset with "
                                  "__len__/__getitem__/get_example_count_per_file (e.g. SyntheticAudioDataset)")


AudioTestingDataset = AudioDataset
