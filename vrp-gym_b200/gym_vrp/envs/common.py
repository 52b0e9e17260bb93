from typing import TypeVar

ObsType = TypeVar("ObsType")
