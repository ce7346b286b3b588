"""Host-side helpers: argument / config printing, attribute dictionaries, checkpoint file readers."""
