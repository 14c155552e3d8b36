"""Drop-in for the reference's ``logger.py`` (Logger.log / write / flush / close, logger.py:6-49) on the stock
``torch.utils.tensorboard`` writer: the reference imports ``tensorboardX``, which this image does not have.
Not on the hot path; kept so that ``train.py:9,31,106-122`` run unchanged.  Scalars are mirrored in memory so
``write()`` can still export ``log.json`` (tensorboardX's ``export_scalars_to_json`` has no stock equivalent)."""
import json
import os
import time

LOG_TYPE = ["scalar", "audio", "image"]


class Logger(object):
    def __init__(self, log_dir):
        from torch.utils.tensorboard import SummaryWriter
        self.log_dir = log_dir
        self.writer = SummaryWriter(log_dir)
        self._scalars = {}

    def log(self, n_iter, report, log_type="scalar", sr=None, text=False):
        if log_type not in LOG_TYPE:
            raise ValueError("Wrong data type for logger.")          # the reference raises a bare string (a TypeError)
        if log_type == "scalar":
            if text:
                self._print_scalars(n_iter, report)
            for k, v in report.items():
                tag = "scalar/{}".format(k)
                self.writer.add_scalar(tag, float(v), n_iter)
                self._scalars.setdefault(os.path.join(self.log_dir, tag), []).append([time.time(), int(n_iter), float(v)])
        elif log_type == "audio":
            if sr is None:
                raise ValueError("Sample rate is required for saving audio data.")
            import torch
            for k, v in report.items():
                self.writer.add_audio(k, torch.as_tensor(v).reshape(1, -1), n_iter, sample_rate=sr)
        else:
            for k, v in report.items():
                self.writer.add_image(k, v, n_iter, dataformats="HWC")   # generate_spec_img returns [H, W, 3]

    def _print_scalars(self, n_iter, report):
        print("---------------------------")
        print("n_iter : {}".format(n_iter))
        for k, v in report.items():
            print("{} : {:.4f}".format(k, v))
        print("---------------------------")

    def write(self):
        os.makedirs(self.log_dir, exist_ok=True)
        with open(os.path.join(self.log_dir, "log.json"), "w") as f:
            json.dump(self._scalars, f)

    def flush(self):
        self.writer.flush()

    def close(self):
        self.writer.close()
