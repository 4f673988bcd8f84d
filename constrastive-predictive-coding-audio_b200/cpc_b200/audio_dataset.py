"""``AudioDataset`` / ``AudioTestingDataset`` (audio_dataset.py:16-199): a folder of audio files seen as one long stream of
samples that is cut into overlapping items (``item_length`` samples every ``unique_length`` samples), with items that
cross file boundaries assembled from consecutive files.  Index arithmetic, item assembly and
``get_example_count_per_file`` (what ``FileBatchSampler`` consumes) follow the reference line by line; only the decoder
differs: the reference calls a torchaudio 0.2-era ``load(normalization=True, num_frames=, offset=)`` that no longer
exists, here PCM / float WAV files are read with the standard library (first channel, integer PCM scaled to [-1, 1) as
``normalization=True`` did) and other containers go through the current ``torchaudio.load`` when it can decode them.
This is SURVEY 8(f) row 4: the host-side input path; decoding is not part of the GPU hot path.
"""
import bisect
import math
import os
import wave
from pathlib import Path

import numpy as np
import torch
import torch.utils.data


def list_all_audio_files(location, allowed_types=(".mp3", ".wav", ".aif", "aiff", ".flac")):
    """audio_dataset.py:266-273: recursive glob per type, each type's matches sorted, types in the given order."""
    files = []
    for suffix in allowed_types:
        files.extend(sorted(Path(location).glob('**/*' + suffix)))
    if not files:
        print("found no audio files in " + str(location))
    return files


def read_wav(path, frames=-1, start=0):
    """First channel of a RIFF WAV file as float32 in [-1, 1): 8 / 16 / 24 / 32-bit PCM (``wave`` cannot open IEEE-float
    files; those go through torchaudio).  ``frames`` = -1 reads to the end."""
    with wave.open(str(path), "rb") as fh:
        channels, width, total = fh.getnchannels(), fh.getsampwidth(), fh.getnframes()
        start = max(0, min(int(start), total))
        count = total - start if frames < 0 else max(0, min(int(frames), total - start))
        fh.setpos(start)
        raw = fh.readframes(count)
    if width == 1:
        data = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    elif width == 2:
        data = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif width == 3:
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        data = (v - ((v & 0x800000) << 1)).astype(np.float32) / 8388608.0
    elif width == 4:
        data = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
    else:
        raise ValueError("unsupported WAV sample width %d in %s" % (width, path))
    return torch.from_numpy(data.reshape(-1, channels)[:, 0].copy())


class AudioDataset(torch.utils.data.Dataset):
    def __init__(self, location, item_length, unique_length=None, sampling_rate=16000, mono=True, dtype=torch.FloatTensor,
                 max_file_count=None, cross_files=True):
        super().__init__()
        self.location = Path(location)
        self.sampling_rate = sampling_rate
        self.mono = mono
        self._item_length = item_length
        self._unique_length = item_length if unique_length is None else unique_length
        self._length = 0
        self.start_samples = [0]
        self.dtype = dtype
        self.dummy_load = False
        self.cross_files = cross_files
        self.files = list_all_audio_files(self.location, allowed_types=['.wav', '.mp3', '.aiff'])
        self.max_file_count = len(self.files) if max_file_count is None else max_file_count
        self.calculate_length()

    @property
    def item_length(self):
        return self._item_length

    @item_length.setter
    def item_length(self, value):
        self._item_length = value
        self.calculate_length()

    @property
    def unique_length(self):
        return self._unique_length

    @unique_length.setter
    def unique_length(self, value):
        self._unique_length = value
        self.calculate_length()

    def load_file(self, file, frames=-1, start=0):
        """(:61-73) first channel, normalised, as ``self.dtype``."""
        if frames == 0:
            print("Error: zero frames requested")
        if os.path.splitext(str(file))[1].lower() == ".wav":
            try:
                return read_wav(file, frames, start).type(self.dtype)
            except wave.Error:
                pass                                             # e.g. IEEE-float WAV: let torchaudio try
        try:
            import torchaudio
            if frames == -1:
                data, _ = torchaudio.load(str(file))
            else:
                data, _ = torchaudio.load(str(file), frame_offset=int(start), num_frames=int(frames))
        except Exception as exc:                                   # noqa: BLE001
            raise RuntimeError("cannot decode %s: only PCM WAV is read natively and torchaudio failed (%s)" % (file, exc))
        return data[0, :].type(self.dtype)

    def file_length(self, path):
        if os.path.splitext(path)[1].lower() == ".wav":
            try:
                with wave.open(path, "rb") as fh:
                    return fh.getnframes()
            except wave.Error:
                pass
        return self.load_file(path).shape[0]

    def calculate_length(self):
        """(:75-95) start sample of every file in the concatenated stream and the number of items."""
        start_samples = [0]
        for idx in range(self.max_file_count):
            next_start = start_samples[-1] + self.file_length(str(self.files[idx]))
            if not self.cross_files:
                next_start -= self.item_length
            start_samples.append(next_start)
        available = start_samples[-1] - (self.item_length - self.unique_length)
        self._length = math.floor(available / self.unique_length)
        self.start_samples = start_samples

    def load_sample(self, file_index, position_in_file, item_length):
        """(:97-113) ``item_length`` (+1 when it fits) samples starting at ``position_in_file``, continuing into the
        following files when the file ends first."""
        file_length = self.start_samples[file_index + 1] - self.start_samples[file_index]
        remaining = position_in_file + item_length + 1 - file_length
        if remaining < 0:
            return self.load_file(str(self.files[file_index]), frames=item_length + 1, start=position_in_file)
        this_part = self.load_file(str(self.files[file_index]), frames=item_length - remaining, start=position_in_file)
        next_part = self.load_sample(file_index + 1, position_in_file=0, item_length=remaining)
        return torch.cat((this_part, next_part))

    def get_position(self, idx):
        """(:115-127) global item index -> (file index, position in that file)."""
        sample_index = idx * self.unique_length
        file_index = bisect.bisect_left(self.start_samples, sample_index) - 1
        if file_index < 0:
            file_index = 0
        if file_index + 1 >= len(self.start_samples):
            print("error: sample index " + str(sample_index) + " is to high. Results in file_index " + str(file_index))
        return file_index, sample_index - self.start_samples[file_index]

    def __getitem__(self, idx):
        if self.dummy_load:
            sample = np.random.randn(self._item_length)
        else:
            file_index, position = self.get_position(idx)
            sample = self.load_sample(file_index, position, self._item_length)
        return sample[:self._item_length]

    def get_segment(self, position, file_index, duration=None):
        """(:139-153)"""
        position_in_file = (position // self.sampling_rate) - self.start_samples[file_index]
        item_length = self._item_length if duration is None else int(duration * self.sampling_rate)
        return self.load_sample(file_index, position_in_file, item_length)

    def get_example_count_per_file(self):
        """(:155-165) items whose first sample lies in each file; what FileBatchSampler groups by."""
        counts = []
        for i in range(1, len(self.start_samples)):
            total = math.ceil(self.start_samples[i] / self.unique_length)
            previous = math.ceil(self.start_samples[i - 1] / self.unique_length)
            counts.append(total - previous)
        surplus = np.sum(counts) - self._length
        if surplus > 0:
            counts[-1] = counts[-1] - surplus
        return counts

    def __len__(self):
        return self._length


class AudioTestingDataset(AudioDataset):
    """(:171-199) items never cross files; ``__getitem__`` also returns the file index (the probe task's label)."""

    def __init__(self, location, item_length, unique_length=None, sampling_rate=16000, mono=True, dtype=torch.FloatTensor,
                 max_file_count=None, cross_files=False):
        super().__init__(location=location, item_length=item_length, unique_length=unique_length,
                         sampling_rate=sampling_rate, mono=mono, dtype=dtype, max_file_count=max_file_count,
                         cross_files=cross_files)

    def __getitem__(self, idx):
        if self.dummy_load:
            sample = np.random.randn(self._item_length)
            file_index = 0
        else:
            file_index, position = self.get_position(idx)
            sample = self.load_sample(file_index, position, self._item_length)
        return sample[:self._item_length], torch.LongTensor([file_index]).squeeze()
