"""CPU restatement of Nextsim::ConstantHealing::updateElement -- TEST INFRASTRUCTURE (oracle).

physics/src/modules/DamageHealingModule/ConstantHealing.cpp:53-72 (default td = 15 days, :13,21-22).
Only tests/ may import this.
"""
import numpy as np


def constant_healing(damage, cice, delta_cice, step_seconds, td_seconds=15 * 86400.0):
    damage = np.array(damage, dtype=np.float64, copy=True)
    lateral = np.maximum(0.0, delta_cice) if delta_cice is not None else np.zeros_like(damage)
    # 1. lateral ice formation: new ice is undamaged (ConstantHealing.cpp:61)
    damage = (damage * (cice - lateral) + lateral) / cice
    # 2. constant (linear) healing (ConstantHealing.cpp:70-71)
    damage += step_seconds / td_seconds
    return np.minimum(1.0, damage)
