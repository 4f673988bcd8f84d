"""Reference module name ``audio_dataset`` (audio_dataset.py:16-273) -> cpc_b200: the batch composition
(FileBatchSampler, on the hot path) and the file-backed datasets (host-side input path)."""
import _bootstrap  # noqa: F401
from cpc_b200.sampler import FileBatchSampler, SyntheticAudioDataset                                 # noqa: F401
from cpc_b200.audio_dataset import AudioDataset, AudioTestingDataset, list_all_audio_files            # noqa: F401
