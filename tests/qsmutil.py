"""Shared by the QSM-algebra tests: rebuild the golden operands (tests/golden/qsmcases.py) for a backend."""

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

import qsmcases  # noqa: E402

GOLD = json.load(open(os.path.join(HERE, "golden", "qsm_vectors.json")))


def oracle_operand(spec):
    from oracle import qsm_np as oq
    return oq.QSM(spec.get("d"), spec.get("lower"), spec.get("upper"), symm=(spec["kind"] == "symm"))


def oracle_operands():
    return {k: oracle_operand(v) for k, v in qsmcases.operands().items()}


TYPE_TO_KIND = {"DiagQSM": "diag", "StrictLowerTriQSM": "strict_lower", "StrictUpperTriQSM": "strict_upper",
                "LowerTriQSM": "lower", "UpperTriQSM": "upper", "SquareQSM": "square", "SymmQSM": "symm"}
