"""Hamiltonian model (drop-in for src/casq/models/hamiltonian_model.py:36-97)."""
from __future__ import annotations

from enum import Enum
from typing import Optional, Union

import numpy as np
import numpy.typing as npt

from .hamiltonian_parser import parse_hamiltonian_dict


class HamiltonianModel:
    """Same constructor and attributes as casq's HamiltonianModel.

    Quirk kept on purpose (hamiltonian_model.py:91-95): the default frame tests the RAW
    ``evaluation_mode`` argument, so ``None`` selects the 1-D diagonal frame while DENSE selects the full
    static operator.
    """

    class EvaluationMode(Enum):
        DENSE = 0
        SPARSE = 1

    def __init__(
        self,
        hamiltonian_dict: dict,
        extracted_qubits: Optional[list[int]] = None,
        rotating_frame: Optional[npt.NDArray] = None,
        in_frame_basis: bool = False,
        evaluation_mode: Optional["HamiltonianModel.EvaluationMode"] = EvaluationMode.DENSE,
        rwa_cutoff_freq: Optional[float] = None,
        rwa_carrier_freqs: Optional[Union[npt.NDArray, tuple[npt.NDArray, npt.NDArray]]] = None,
    ) -> None:
        self.hamiltonian_dict = hamiltonian_dict
        self.in_frame_basis = in_frame_basis
        self.rwa_carrier_freqs = rwa_carrier_freqs
        self.extracted_qubits = [0] if extracted_qubits is None else extracted_qubits
        self.evaluation_mode = HamiltonianModel.EvaluationMode.DENSE if evaluation_mode is None else evaluation_mode
        self.rwa_cutoff_freq = rwa_cutoff_freq
        (self.static_operator, self.operators, self.channels, self.qubit_dims) = parse_hamiltonian_dict(
            hamiltonian_dict, extracted_qubits
        )
        if rotating_frame is None:
            if evaluation_mode is HamiltonianModel.EvaluationMode.DENSE:
                self.rotating_frame = self.static_operator
            else:
                self.rotating_frame = np.diag(self.static_operator)
        else:
            self.rotating_frame = rotating_frame
