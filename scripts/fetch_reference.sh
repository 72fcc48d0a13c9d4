#!/usr/bin/env bash
# Attempt to install the UNMODIFIED reference (sinel/casq) and its pinned arithmetic dependencies (qiskit 0.43.3,
# qiskit-dynamics 0.4.1, jax[cpu] 0.4.6, scipy 1.9.3 -- poetry.lock of the reference) into baseline/_ref so that the
# oracle could be pinned against real QiskitPulseBackend.solve output.  No network: only the offline wheelhouse can
# serve.  The outcome (success or refusal) is logged to profiles/r02_reference_install.log and committed.
set -u
cd "$(dirname "$0")/.."
LOG=profiles/r02_reference_install.log
{
  echo "# $(date -u +%FT%TZ) host=$(hostname) python=$(python -V 2>&1)"
  echo "## wheelhouse entries matching qiskit|jax|dynamics|symengine|rustworkx|matplotlib|qutip|multiset:"
  ls /opt/wheelhouse 2>/dev/null | grep -i -E "qiskit|jax|dynamics|symengine|rustworkx|matplotlib|qutip|multiset" || echo "(none)"
  echo "## importable now?"
  for m in qiskit qiskit_dynamics jax matplotlib symengine; do python -c "import $m" 2>&1 | tail -1 | sed "s/^/$m: /"; python -c "import $m" 2>/dev/null && echo "$m: importable"; done
  SRC=/root/reference
  if [ -d "$SRC" ]; then
    echo "## pip install (with dependency resolution) from the offline wheelhouse"
    python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref "$SRC" 2>&1 | tail -15
    echo "## pip install --no-deps (the package alone; its imports still need qiskit)"
    rm -rf /tmp/casq_src && cp -r "$SRC" /tmp/casq_src
    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref /tmp/casq_src 2>&1 | tail -8
    echo "## import casq from baseline/_ref"
    PYTHONPATH=baseline/_ref python -c "import casq.backends" 2>&1 | tail -3
  else
    echo "## /root/reference absent on this machine (GPU box): pinned wheels only"
    python -m pip download --no-deps -d /tmp/whl qiskit-dynamics==0.4.1 2>&1 | tail -3
  fi
} > "$LOG" 2>&1
cat "$LOG"
