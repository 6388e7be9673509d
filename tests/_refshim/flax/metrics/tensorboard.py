class SummaryWriter:
    def __init__(self, *a, **k):
        self.scalars = []

    def scalar(self, name, value, step):
        self.scalars.append((name, float(value), int(step)))

    def image(self, *a, **k):
        pass
