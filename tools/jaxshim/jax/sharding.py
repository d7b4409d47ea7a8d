class Mesh:
    def __init__(self, devices, axis_names):
        self.devices, self.axis_names = devices, axis_names


class PartitionSpec(tuple):
    def __new__(cls, *a):
        return super().__new__(cls, a)


class NamedSharding:
    def __init__(self, mesh, spec):
        self.mesh, self.spec = mesh, spec
