"""Wall-clock timer that prints the reference's ``[timer]`` lines.

The line format is the one distllm/timer.py:156-162 emits and ``TimeLogger.parse_logs``
(distllm/timer.py:129-154) reads back, because those lines are how embed runs are timed:

    [timer] [<tags>] in [<seconds>] seconds.  start: [<unix>], end: [<unix>]
"""

from __future__ import annotations

import re
import time
from pathlib import Path
from typing import Any
from typing import NamedTuple
from typing import Sequence
from typing import Union

PathLike = Union[Path, str]
_BRACKETED = re.compile(r'\[([^\[\]]+)\]')


class TimeStats(NamedTuple):
    tags: Sequence[str]
    elapsed_s: float
    start_unix: float
    end_unix: float


class Timer:
    """Context manager / start-stop timer; logs one ``[timer]`` line when stopped."""

    def __init__(self, *tags: Any) -> None:
        self.tags = tags
        self._t0 = 0
        self._t1 = 0
        self._unix0 = 0.0
        self._unix1 = 0.0
        self._running = False

    def start(self) -> 'Timer':
        self._running = True
        self._unix0 = time.time()
        self._t0 = time.perf_counter_ns()
        return self

    def stop(self) -> None:
        self._t1 = time.perf_counter_ns()
        self._unix1 = time.time()
        self._running = False
        TimeLogger().log(TimeStats(self.tags, self.elapsed_s, self._unix0, self._unix1))

    def __enter__(self) -> 'Timer':
        return self.start()

    def __exit__(self, *exc: object) -> None:
        self.stop()

    @property
    def elapsed_ns(self) -> int:
        if self._running:
            raise RuntimeError('Timer is still running!')
        return self._t1 - self._t0

    @property
    def elapsed_ms(self) -> float:
        return self.elapsed_ns / 1e6

    @property
    def elapsed_s(self) -> float:
        return self.elapsed_ns / 1e9


class TimeLogger:
    """Emit and parse ``[timer]`` lines."""

    def log(self, ts: TimeStats) -> None:
        tags = ' '.join(str(t) for t in ts.tags)
        print(
            f'[timer] [{tags}] in [{ts.elapsed_s:.2f}] seconds.',
            f' start: [{ts.start_unix:.2f}], end: [{ts.end_unix:.2f}]',
            flush=True,
        )

    def parse_logs(self, log_path: PathLike) -> list[TimeStats]:
        out = []
        for line in Path(log_path).read_text().splitlines():
            if '[timer]' not in line:
                continue
            fields = _BRACKETED.findall(line)
            # fields: 'timer', tags, seconds, start, end (kept as strings like the reference)
            out.append(TimeStats(fields[1].split(), fields[2], fields[3], fields[4]))
        return out
