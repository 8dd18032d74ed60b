from .. import constant_init, kaiming_init, xavier_init, normal_init  # noqa
