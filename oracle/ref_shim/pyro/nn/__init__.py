import torch


class DenseNN(torch.nn.Module):
    """pyro.nn.DenseNN (pyro-ppl 1.8.6) semantics: Linear -> f -> ... -> Linear, no activation on the output;
    returns a tensor when len(param_dims) == 1, else a tuple of slices."""

    def __init__(self, input_dim, hidden_dims, param_dims=[1, 1], nonlinearity=torch.nn.ReLU()):
        super().__init__()
        self.input_dim = input_dim
        self.hidden_dims = hidden_dims
        self.param_dims = param_dims
        self.count_params = len(param_dims)
        self.output_multiplier = sum(param_dims)
        ends = torch.cumsum(torch.tensor(param_dims), dim=0)
        starts = torch.cat((torch.zeros(1).type_as(ends), ends[:-1]))
        self.param_slices = [slice(s.item(), e.item()) for s, e in zip(starts, ends)]
        layers = [torch.nn.Linear(input_dim, hidden_dims[0])]
        for i in range(1, len(hidden_dims)):
            layers.append(torch.nn.Linear(hidden_dims[i - 1], hidden_dims[i]))
        layers.append(torch.nn.Linear(hidden_dims[-1], self.output_multiplier))
        self.layers = torch.nn.ModuleList(layers)
        self.f = nonlinearity

    def forward(self, x):
        h = x
        for layer in self.layers[:-1]:
            h = self.f(layer(h))
        h = self.layers[-1](h)
        if self.output_multiplier == 1:
            return h
        h = h.reshape(list(x.size()[:-1]) + [self.output_multiplier])
        if self.count_params == 1:
            return h
        return tuple(h[..., s] for s in self.param_slices)
