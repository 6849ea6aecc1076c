from openscene_b200.resnet_block import BasicBlock, Bottleneck  # noqa: F401
