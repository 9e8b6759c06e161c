"""Pruner registry: the slice of lavis/common/registry.py (:113-137, :269-270) the hot path uses."""


class Registry:
    mapping = {"pruner_name_mapping": {}}

    @classmethod
    def register_pruner(cls, name):
        def wrap(pruner_cls):
            if name in cls.mapping["pruner_name_mapping"]:
                raise KeyError(f"Name '{name}' already registered for "
                               f"{cls.mapping['pruner_name_mapping'][name]}.")
            cls.mapping["pruner_name_mapping"][name] = pruner_cls
            return pruner_cls
        return wrap

    @classmethod
    def get_pruner_class(cls, name):
        return cls.mapping["pruner_name_mapping"].get(name, None)

    @classmethod
    def list_pruners(cls):
        return sorted(cls.mapping["pruner_name_mapping"].keys())


registry = Registry()
