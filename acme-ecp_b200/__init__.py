"""acme-ecp_b200: B200-native (sm_100a) implementation of SP-E3SM's embedded cloud-resolving model behind the
reference's crm() argument surface (E3SM-Project/ACME-ECP, components/cam/src/physics/crm/crm_module.F90:45).

The package is a thin host layer over the C-ABI shared library built from csrc/ (include/crm_b200.h).  There is
no CPU fallback: importing .crm and calling crm() without the CUDA library raises.
"""
from .crm_types import (CrmDims, crm_input_type, crm_state_type, crm_rad_type, crm_output_type,
                        crm_ecpp_output_type)
from . import abi, synthetic, shard, replay

__all__ = ["CrmDims", "crm_input_type", "crm_state_type", "crm_rad_type", "crm_output_type",
           "crm_ecpp_output_type", "abi", "synthetic", "shard", "replay"]
