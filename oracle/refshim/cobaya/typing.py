from typing import Any, Dict

empty_dict: Dict[str, Any] = {}
TheoryDictIn = Dict[str, Any]
