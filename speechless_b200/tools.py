"""Small helpers the hot path uses; same names and behaviour as the reference's
`speechless/tools.py` (`single` :15-20, `single_or_none` :23-26, `average_or_nan` :91-95,
`paginate` :98-100, `mkdir` :44-48, `log`/`logger` :103-112)."""
import logging
import sys
from os import makedirs
from pathlib import Path
from typing import Any, Iterable, List, Optional, TypeVar

E = TypeVar("E")


def single(sequence: List[E]) -> E:
    if len(sequence) != 1:
        raise AssertionError("expected exactly one element, got {}".format(len(sequence)))
    return sequence[0]


def single_or_none(sequence: List[E]) -> Optional[E]:
    if len(sequence) > 1:
        raise AssertionError("expected at most one element, got {}".format(len(sequence)))
    return sequence[0] if sequence else None


def average_or_nan(numbers: List[float]) -> float:
    return sum(numbers) / len(numbers) if len(numbers) else float("nan")


def paginate(sequence: List[E], page_size: int) -> Iterable[List[E]]:
    for start in range(0, len(sequence), page_size):
        yield sequence[start:start + page_size]


def mkdir(directory: Path) -> None:
    makedirs(str(directory), exist_ok=True)


def read_text(path: Path, encoding=None) -> str:
    with Path(path).open(encoding=encoding) as f:
        return f.read()


logger = logging.getLogger("results")
logger.setLevel(logging.INFO)
if not logger.handlers:
    _handler = logging.StreamHandler(sys.stdout)
    _handler.setLevel(logging.INFO)
    logger.addHandler(_handler)


def log(obj: Any) -> None:
    logger.info(str(obj))
