from openscene_b200.me_utils import kaiming_normal_  # noqa: F401
