"""Drop-in for the reference's ``dcn`` package (dcn/__init__.py:1-7): same public names."""
from .deform_conv import (DeformConv, DeformConvPack, ModulatedDeformConv, ModulatedDeformConvPack,
                          ModulatedDeformConvPack2, deform_conv, modulated_deform_conv)

__all__ = [
    'DeformConv', 'DeformConvPack', 'ModulatedDeformConv', 'ModulatedDeformConvPack',
    'ModulatedDeformConvPack2', 'deform_conv', 'modulated_deform_conv'
]
