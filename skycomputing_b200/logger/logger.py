"""Plain file logger (line format identical to scaelum/logger/logger.py:4-14)."""
from __future__ import annotations

import os
from datetime import datetime


class Logger:
    def __init__(self, filename: str, mode: str = "a"):
        d = os.path.dirname(os.path.abspath(filename))
        os.makedirs(d, exist_ok=True)
        self.filename = filename
        self.file = open(file=filename, mode=mode)

    def info(self, message: str) -> None:
        self._write("INFO - {} - {}\n".format(datetime.now(), message))

    def _write(self, message: str) -> None:
        self.file.write(message)
        self.file.flush()

    def close(self) -> None:
        try:
            self.file.close()
        except Exception:
            pass

    def __del__(self):
        self.close()
