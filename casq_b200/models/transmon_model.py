"""Transmon (Duffing oscillator) model (drop-in for src/casq/models/transmon_model.py:35-170)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Union

import numpy.typing as npt

from .hamiltonian_model import HamiltonianModel

# ibmq_manila constants (transmon_model.py:94-114); rad/s
_MANILA_WQ = (31179079102.853794, 30397743782.610542, 31649945798.50227, 31107813662.24873, 31825180853.3539)
_MANILA_DELTA = (-2165345334.8252344, -2169482392.6367006, -2152313197.3287387, -2158766696.6684937, -2149525690.7311115)
_MANILA_OMEGAD = (926545606.6640488, 892870223.8110852, 927794953.0001632, 921439621.8693779, 1150709205.1097605)
_MANILA_J = {(0, 1): 11845444.218797993, (1, 2): 11967839.68906386, (2, 3): 12402113.956012368, (3, 4): 12186910.37040823}
# (amplitude qubit, operator qubit, u-channel index) of the cross-drive terms, transmon_model.py:82-89
_MANILA_U = ((1, 0, 0), (0, 1, 1), (2, 1, 2), (1, 2, 3), (3, 2, 4), (4, 3, 6), (2, 3, 5), (3, 4, 7))


def _manila() -> dict:
    h_str = [
        "_SUM[i,0,4,wq{i}/2*(I{i}-Z{i})]",
        "_SUM[i,0,4,delta{i}/2*O{i}*O{i}]",
        "_SUM[i,0,4,-delta{i}/2*O{i}]",
        "_SUM[i,0,4,omegad{i}*X{i}||D{i}]",
    ]
    variables = {}
    for (a, b), j in _MANILA_J.items():
        h_str += [f"jq{a}q{b}*Sp{a}*Sm{b}", f"jq{a}q{b}*Sm{a}*Sp{b}"]
        variables[f"jq{a}q{b}"] = j
    h_str += [f"omegad{amp}*X{op}||U{u}" for amp, op, u in _MANILA_U]
    for q in range(5):
        variables[f"wq{q}"] = _MANILA_WQ[q]
        variables[f"delta{q}"] = _MANILA_DELTA[q]
        variables[f"omegad{q}"] = _MANILA_OMEGAD[q]
    return {
        "description": "ibmq_manila: five Duffing-oscillator transmons with exchange coupling (rad/s).",
        "h_str": h_str,
        "osc": {},
        "qub": {str(q): 3 for q in range(5)},
        "vars": variables,
    }


class TransmonModel(HamiltonianModel):
    """wq*N + delta/2*N(N-1) + omegad*X*D(t) per qubit, exchange coupling, cross-drive U terms; 3 levels."""

    MANILA = _manila()

    @dataclass
    class TransmonProperties:
        frequency: float
        anharmonicity: float
        drive: float

    def __init__(
        self,
        qubit_map: dict[int, "TransmonModel.TransmonProperties"],
        coupling_map: dict[tuple[int, int], float],
        extracted_qubits: Optional[list[int]] = None,
        rotating_frame: Optional[npt.NDArray] = None,
        in_frame_basis: bool = False,
        evaluation_mode: HamiltonianModel.EvaluationMode = HamiltonianModel.EvaluationMode.DENSE,
        rwa_cutoff_freq: Optional[float] = None,
        rwa_carrier_freqs: Optional[Union[npt.NDArray, tuple[npt.NDArray, npt.NDArray]]] = None,
    ) -> None:
        q_max = len(qubit_map.keys()) - 1
        h_str = [
            f"_SUM[i,0,{q_max},wq{{i}}/2*(I{{i}}-Z{{i}})]",
            f"_SUM[i,0,{q_max},delta{{i}}/2*O{{i}}*O{{i}}]",
            f"_SUM[i,0,{q_max},-delta{{i}}/2*O{{i}}]",
            f"_SUM[i,0,{q_max},omegad{{i}}*X{{i}}||D{{i}}]",
        ]
        h_qub, h_vars = {}, {}
        for q, props in qubit_map.items():
            h_qub[str(q)] = 3
            h_vars[f"wq{q}"] = props.frequency
            h_vars[f"delta{q}"] = props.anharmonicity
            h_vars[f"omegad{q}"] = props.drive
        for (q0, q1), strength in coupling_map.items():
            h_vars[f"jq{q0}q{q1}"] = strength
            h_str.append(f"jq{q0}q{q1}*Sp{q0}*Sm{q1}")
            h_str.append(f"jq{q0}q{q1}*Sp{q1}*Sm{q0}")
            h_str.append(f"omegad{q1}*X{q0}||U{q0}")
            h_str.append(f"omegad{q0}*X{q1}||U{q1}")
        super().__init__(
            hamiltonian_dict={"h_str": h_str, "qub": h_qub, "vars": h_vars},
            extracted_qubits=extracted_qubits,
            rotating_frame=rotating_frame,
            in_frame_basis=in_frame_basis,
            evaluation_mode=evaluation_mode,
            rwa_cutoff_freq=rwa_cutoff_freq,
            rwa_carrier_freqs=rwa_carrier_freqs,
        )

    @classmethod
    def manila(cls, extracted_qubits: Optional[list[int]] = None, **kwargs) -> HamiltonianModel:
        """HamiltonianModel over the full MANILA dictionary (what ``from_backend(FakeManila())`` would parse)."""
        return HamiltonianModel(hamiltonian_dict=cls.MANILA, extracted_qubits=extracted_qubits, **kwargs)
