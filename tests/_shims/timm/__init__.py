"""Four-symbol stand-in for `timm`, only so that the reference (`/root/reference`, which imports
`timm.layers.{DropPath,to_2tuple,to_3tuple}` and `timm.models.vision_transformer.trunc_normal_`) can
be imported in the build container to GENERATE golden fixtures (tests/golden/make_golden.py).  It is
never imported by the product (aurora_b200/) nor by anything that runs on the GPU box."""
